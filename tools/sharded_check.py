"""Multi-GPU check of the Gaussian-sharded render (run under torchrun on >= 2 GPUs):
the sharded result must equal the single-GPU result of the same cloud (bit-identical images, gradients to 1e-4),
and the script prints the fwd+bwd time of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py [--workload C4]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "4dgs-slam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--exchange", default="alltoall,allgather")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import sharded
    from tools import runners
    from tools.scenes import broadcast_scene, config_scene, make_scene

    sc_cpu = config_scene(args.workload) if args.workload.startswith("C") else make_scene(20000, 320, 240, sh_degree=2, seed=5)
    sc = broadcast_scene(sc_cpu.to(dev))       # all ranks render rank 0's (bit-identical) scene
    rs = runners.settings_for(sc, dgr)
    lo, hi = sharded.shard_bounds(sc.P, world, rank)

    def run_sharded(exchange="alltoall"):
        leaf = {k: getattr(sc, k)[lo:hi].detach().clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta, rho = torch.zeros(3, device=dev, requires_grad=True), torch.zeros(3, device=dev, requires_grad=True)
        r = sharded.ShardedGaussianRasterizer(rs, exchange=exchange)
        color, radii, depth, opacity, n_touched = r(means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"],
                                                    scales=leaf["scales"], rotations=leaf["rotations"], theta=theta, rho=rho)
        ((color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()).backward()
        return dict(color=color.detach(), depth=depth.detach(), opacity=opacity.detach(), radii=radii, n_touched=n_touched,
                    dL_dmeans3D=leaf["means3D"].grad, dL_dscales=leaf["scales"].grad, dL_drots=leaf["rotations"].grad,
                    dL_dopacity=leaf["opacities"].grad, dL_dshs=leaf["shs"].grad, dL_dmeans2D=m2d.grad,
                    dL_dtau=torch.cat([rho.grad.reshape(-1), theta.grad.reshape(-1)]))

    def timeit(fn, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    single = runners.run_public_api(sc, dgr)           # the whole cloud on this GPU
    rep = {}
    for ex in args.exchange.split(","):
        out = run_sharded(ex)
        for k in ("color", "depth", "opacity"):
            rep[f"{ex}_{k}_bit_identical"] = bool(torch.equal(out[k], single[k]))
        rep[f"{ex}_radii_equal"] = bool(torch.equal(out["radii"], single["radii"][lo:hi]))
        rep[f"{ex}_n_touched_equal"] = bool(torch.equal(out["n_touched"], single["n_touched"][lo:hi]))
        for k in ("dL_dmeans3D", "dL_dscales", "dL_drots", "dL_dopacity", "dL_dshs", "dL_dmeans2D"):
            a, b = out[k].double().reshape(-1), single[k][lo:hi].double().reshape(-1)
            rep[f"{ex}_{k}_l2_rel"] = float((a - b).norm() / b.norm())
        a, b = out["dL_dtau"].double(), single["dL_dtau"].double()
        rep[f"{ex}_dL_dtau_l2_rel"] = float((a - b).norm() / b.norm())
        rep[f"ms_fwd_bwd_sharded_{ex}"] = timeit(lambda: run_sharded(ex))
    ms_single = timeit(lambda: runners.run_public_api(sc, dgr))
    rep.update(workload=args.workload, world=world, P=sc.P, ms_fwd_bwd_single_gpu=ms_single)
    gathered = [None] * world
    dist.all_gather_object(gathered, rep)
    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        with open(os.path.join(args.out, f"sharded_check_{args.workload}_x{world}.json"), "w") as f:
            json.dump(gathered, f, indent=1)
        print(json.dumps(gathered[0]))
        ok = all(all(v for k, v in g.items() if k.endswith(("identical", "equal"))) and all(v < 1e-4 for k, v in g.items() if k.endswith("l2_rel"))
                 for g in gathered)
        print("SHARDED CHECK", "PASS" if ok else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
