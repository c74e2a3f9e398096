"""Where the time of a Gaussian-sharded frame goes (run under torchrun): host enqueue time vs device time, and the device
timeline of rank 0 (kernels + NCCL) from torch.profiler.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/sharded_profile.py --workload X4
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "4dgs-slam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="X4")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import sharded
    from tools import runners
    from tools.sharded_check import load_scene
    sc = load_scene(args.workload, dev)
    rs = runners.settings_for(sc, dgr)
    lo, hi = sharded.shard_bounds(sc.P, world, rank)
    shard = {k: getattr(sc, k)[lo:hi].detach().clone() for k in ("means3D", "opacities", "shs", "scales", "rotations")}

    def fwd():
        leaf = {k: v.detach().requires_grad_(True) for k, v in shard.items()}
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta, rho = torch.zeros(3, device=dev, requires_grad=True), torch.zeros(3, device=dev, requires_grad=True)
        r = sharded.ShardedGaussianRasterizer(rs)
        color, radii, depth, opacity, n_touched = r(means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"],
                                                    scales=leaf["scales"], rotations=leaf["rotations"], theta=theta, rho=rho)
        return (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()

    for _ in range(3):
        fwd().backward()
    torch.cuda.synchronize()
    dist.barrier(device_ids=[local])
    rows = []
    for _ in range(args.iters):
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = fwd()
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        rows.append((t1 - t0, t2 - t1, t3 - t0))
    med = lambda i: sorted(r[i] for r in rows)[len(rows) // 2] * 1e3
    rep = dict(rank=rank, host_forward_ms=med(0), host_backward_ms=med(1), wall_fwd_bwd_ms=med(2))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            fwd().backward()
        torch.cuda.synchronize()
    ev = prof.key_averages()
    top = sorted(ev, key=lambda e: -e.device_time_total)[:22]
    rep["device_us_per_frame"] = {e.key[:70]: round(e.device_time_total / 3, 1) for e in top if e.device_time_total > 0}
    topc = sorted(ev, key=lambda e: -e.self_cpu_time_total)[:14]
    rep["host_self_us_per_frame"] = {e.key[:70]: round(e.self_cpu_time_total / 3, 1) for e in topc}
    out = [None] * world
    dist.all_gather_object(out, rep)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"sharded_profile_{args.workload}_x{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out[0], indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
