"""Multi-GPU check + timing of the Gaussian-sharded render (run under torchrun on >= 2 GPUs; also imported by bench.py):
the sharded result must equal the single-GPU result of the same cloud (bit-identical images, gradients to 1e-4) -- and, for the
bit-reproducible X scenes, the committed digests of the UNMODIFIED reference build -- and the fwd+bwd time of both is reported.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py [--workload X4]
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


def load_scene(workload, dev):
    """X* = tools.scenes.exact_scene (bit-reproducible on every rank); C* / small = rank 0's scene broadcast to all."""
    from tools.scenes import EXACT_CONFIGS, LARGE_CONFIGS, broadcast_scene, config_scene, exact_scene, make_scene
    if workload in EXACT_CONFIGS or workload in LARGE_CONFIGS:
        return exact_scene(workload).to(dev)
    sc_cpu = config_scene(workload) if workload.startswith("C") else make_scene(20000, 320, 240, sh_degree=2, seed=5)
    return broadcast_scene(sc_cpu.to(dev))


def measure(workload, iters, dev, rank, world, local, group=None, single_gpu=True):
    """Returns this rank's report: parity flags of the sharded render against one GPU (and the reference digests when they
    exist), max-over-ranks fwd+bwd milliseconds of the sharded render, and the single-GPU milliseconds of the same frame."""
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import sharded
    from tools import digests, runners

    sc = load_scene(workload, dev)
    rs = runners.settings_for(sc, dgr)
    lo, hi = sharded.shard_bounds(sc.P, world, rank)
    shard = {k: getattr(sc, k)[lo:hi].detach().clone() for k in ("means3D", "opacities", "shs", "scales", "rotations")}

    def run_sharded(deferred=False):
        leaf = {k: v.requires_grad_(True) for k, v in ((k, v.detach()) for k, v in shard.items())}
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta, rho = torch.zeros(3, device=dev, requires_grad=True), torch.zeros(3, device=dev, requires_grad=True)
        r = sharded.ShardedGaussianRasterizer(rs, group=group, deferred_check=deferred)
        color, radii, depth, opacity, n_touched = r(means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"],
                                                    scales=leaf["scales"], rotations=leaf["rotations"], theta=theta, rho=rho)
        ((color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()).backward()
        return dict(color=color.detach(), depth=depth.detach(), opacity=opacity.detach(), radii=radii, n_touched=n_touched,
                    dL_dmeans3D=leaf["means3D"].grad, dL_dscales=leaf["scales"].grad, dL_drots=leaf["rotations"].grad,
                    dL_dopacity=leaf["opacities"].grad, dL_dshs=leaf["shs"].grad, dL_dmeans2D=m2d.grad,
                    dL_dtau=torch.cat([rho.grad.reshape(-1), theta.grad.reshape(-1)]))

    def timeit(fn, warmup=3, collective=True):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if collective:
            dist.barrier(group=group, device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        if collective:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    rep = dict(workload=workload, world=world, P=sc.P, W=sc.W, H=sc.H)
    out = run_sharded()
    if single_gpu:
        single = runners.run_public_api(sc, dgr)           # the whole cloud on this GPU
        for k in ("color", "depth", "opacity"):
            rep[f"{k}_bit_identical_to_1gpu"] = bool(torch.equal(out[k], single[k]))
        rep["radii_equal"] = bool(torch.equal(out["radii"], single["radii"][lo:hi]))
        rep["n_touched_equal"] = bool(torch.equal(out["n_touched"], single["n_touched"][lo:hi]))
        for k in ("dL_dmeans3D", "dL_dscales", "dL_drots", "dL_dopacity", "dL_dshs", "dL_dmeans2D"):
            a, b = out[k].double().reshape(-1), single[k][lo:hi].double().reshape(-1)
            rep[f"{k}_l2_rel"] = float((a - b).norm() / b.norm())
        a, b = out["dL_dtau"].double(), single["dL_dtau"].double()
        rep["dL_dtau_l2_rel"] = float((a - b).norm() / b.norm())
        del single
    dpath = os.path.join(ROOT, "tests", "golden", "digests_ref.json")
    if os.path.exists(dpath):
        ref = json.load(open(dpath)).get(workload)
        if ref is not None:
            import hashlib
            import numpy as np
            for k in ("color", "depth", "opacity"):
                a = np.ascontiguousarray(out[k].cpu().numpy().astype(np.float32).reshape(-1))
                rep[f"{k}_matches_reference_digest"] = hashlib.sha256(a.tobytes()).hexdigest() == ref[k]["sha256"]
    rep["ms_fwd_bwd_sharded"] = timeit(run_sharded)
    # deferred capacity check (the host never waits on the device inside a frame); same kernels, same results
    out_d = run_sharded(True)
    rep["deferred_color_bit_identical"] = bool(torch.equal(out_d["color"], out["color"]))
    rep["ms_fwd_bwd_sharded_deferred_check"] = timeit(lambda: run_sharded(True))
    cap = sharded.last_capacity(dev, sc.W, sc.H, world)
    rep["slab_capacity_per_pair"] = cap
    rep["bytes_sent_per_rank"] = sharded.collective_bytes(sc.W, sc.H, world, cap)
    rep["redos"] = sharded._state((str(dev), sc.W, sc.H, world))["redos"]
    if single_gpu:
        rep["ms_fwd_bwd_single_gpu"] = timeit(lambda: runners.run_public_api(sc, dgr), collective=False)
        rep["speedup_vs_single_gpu"] = rep["ms_fwd_bwd_single_gpu"] / rep["ms_fwd_bwd_sharded"]
        rep["speedup_vs_single_gpu_deferred_check"] = rep["ms_fwd_bwd_single_gpu"] / rep["ms_fwd_bwd_sharded_deferred_check"]
    return rep


def passed(rep: dict) -> bool:
    flags = all(v for k, v in rep.items() if k.endswith(("identical_to_1gpu", "equal", "reference_digest", "bit_identical")))
    return flags and all(v < 1e-4 for k, v in rep.items() if k.endswith("l2_rel"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="X2")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rep = measure(args.workload, args.iters, dev, rank, world, local)
    gathered = [None] * world
    dist.all_gather_object(gathered, rep)
    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        with open(os.path.join(args.out, f"sharded_check_{args.workload}_x{world}.json"), "w") as f:
            json.dump(gathered, f, indent=1)
        print(json.dumps(gathered[0]))
        print("SHARDED CHECK", "PASS" if all(passed(g) for g in gathered) else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
