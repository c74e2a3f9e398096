#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (BASELINE.json metric: rasterizer fwd+bwd frames/sec @640x480, 500k Gaussians).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3|C2|C3sh3|C4|C3map|track]

A *step* is one forward+backward pass of the differentiable rasterizer over one synthetic frame of the workload
(tools/scenes.py; data = synthetic, seeded).  One JSON line is printed by rank 0:

  value        frames/s with every input already resident in HBM, through the public GaussianRasterizer API
  e2e          the same metric with HOST inputs: every step copies the frame's Gaussian tensors + camera from pinned host
               memory to the device (double-buffered, overlapped with the previous step's compute), runs fwd+bwd through the
               public API and reads the loss and the pose gradient back; one timed region around all K steps
  roofline     dominant kernel: algorithmic bytes (SURVEY.md 8d terms) / its mean duration measured live with CUDA events
               on the launching stream (g4r_profile_*), against MEASURED_PEAKS.json's HBM copy bandwidth
  roofline_frame  whole-frame B_alg / t_step (the "fraction of HBM roofline" of BASELINE.md section 2d)
  cpu_baseline the CPU oracle (a port: the reference has no CPU rasterizer) on the box's host cores, bounded sample
  --impl reference  times the UNMODIFIED reference CUDA build (baseline/_ref) through its own public API on the same
               workload; if that build is absent it times the CPU oracle port instead and says so.

N > 1 (torchrun, one rank per GPU): the path shards over independent views of the same cloud (the mapping loop renders
~10 keyframes per iteration); every rank renders its own view, no data-path collective, weak scaling; time = max over
ranks of the device time.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "4dgs-slam_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self, t0: float = None, t1: float = None) -> dict:
        """Samples whose timestamp falls inside the timed region [t0, t1] (host wall clock); all samples if none do."""
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        use = inside if inside else rows
        reasons = set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in use]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in use) if use else None,
                "reasons": sorted(reasons), "samples": len(use), "samples_inside_timed_region": len(inside)}


def algorithmic_bytes(P, K, N, X, tiles):
    """SURVEY.md section 8d / BASELINE.md section 2d.  S = reference sort passes over (32 + bit) key bits."""
    # getHigherMsb (rasterizer_impl.cu:35-50) returns the position above the MSB: 9/11/13 for 300/1200/4800 tiles
    bit = int(tiles).bit_length()
    S = math.ceil((32 + bit) / 8)
    b_fwd = P * (44 + 12 * K) + 56 * P + N * (12 + 24 * S + 48) + 28 * X
    b_bwd = N * (48 + 40) + 24 * X + P * (44 + 12 * K) + 88 * P + P * (56 + 12 * K)
    per_kernel = {
        "project": P * (44 + 12 * K) + 56 * P,
        "binning": N * (12 + 24 * S),
        "composite_forward": N * 48 + 28 * X,
        "composite_backward": N * (48 + 40) + 24 * X,
        "gaussian_backward": P * (44 + 12 * K) + 88 * P + P * (56 + 12 * K),
    }
    return b_fwd + b_bwd, per_kernel, S


# ---------------------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, name, seed, device):
        from tools.scenes import config_scene
        self.cpu = config_scene(name, seed=seed)
        self.name = name
        self.dev = self.cpu.to(device)
        self.device = device
        # pinned host copies for the end-to-end leg
        # (one pinned staging buffer, every input a 256-byte-aligned view of it: a step's inputs then travel as ONE copy instead
        # of a dozen small cudaMemcpyAsync calls -- the e2e loop is host-bound before it is PCIe-bound)
        src = {}
        for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp", "viewmatrix", "projmatrix",
                  "projmatrix_raw", "campos", "bg"):
            v = getattr(self.cpu, k)
            if v is not None:
                src[k] = v.contiguous()
        self.layout, off = {}, 0
        for k, v in src.items():
            self.layout[k] = (off, v.numel() * v.element_size(), v.dtype, tuple(v.shape))
            off += (v.numel() * v.element_size() + 255) // 256 * 256
        self.host_pack = torch.zeros(off, dtype=torch.uint8).pin_memory()
        self.host = self.views(self.host_pack)
        for k, v in src.items():
            self.host[k].copy_(v)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.d2h_bytes = 4 + 6 * 4                       # loss scalar + pose gradient (rho, theta)
        self.grad_color, self.grad_depth = self.dev.grad_color, self.dev.grad_depth
        self.result_host = torch.empty(7, dtype=torch.float32).pin_memory()


def _views(self, pack):
    """The input tensors as views of a packed uint8 buffer (host or device) laid out like self.host_pack."""
    return {k: pack[o:o + n].view(dt).view(shape) for k, (o, n, dt, shape) in self.layout.items()}


Workload.views = _views


def make_step(dgr, wl: Workload, from_host: bool):
    """Returns fn() -> None doing one fwd+bwd through the public API of module `dgr`."""
    sc = wl.dev
    dev = wl.device

    def settings(t):
        return dgr.GaussianRasterizationSettings(image_height=sc.H, image_width=sc.W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=t["bg"],
                                                 scale_modifier=sc.scale_modifier, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"],
                                                 projmatrix_raw=t["projmatrix_raw"], sh_degree=sc.sh_degree, campos=t["campos"],
                                                 prefiltered=False, debug=False)

    resident = {k: getattr(sc, k) for k in wl.host}
    leaf_keys = [k for k in ("means3D", "opacities", "shs", "scales", "rotations") if k in wl.host]

    def step():
        t = {k: v.to(dev, non_blocking=True) for k, v in wl.host.items()} if from_host else resident
        leaf = {k: t[k].detach().requires_grad_(True) for k in leaf_keys}
        means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(settings(t))(
            means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=leaf.get("shs"),
            colors_precomp=t.get("colors_precomp"), scales=leaf.get("scales"), rotations=leaf.get("rotations"),
            cov3D_precomp=t.get("cov3D_precomp"), theta=theta, rho=rho)
        loss = (color * wl.grad_color).sum() + (depth * wl.grad_depth).sum()
        loss.backward()
        if from_host:
            wl.result_host[:1].copy_(loss.detach().reshape(1), non_blocking=True)
            wl.result_host[1:4].copy_(rho.grad, non_blocking=True)
            wl.result_host[4:7].copy_(theta.grad, non_blocking=True)
    return step


def run_e2e(dgr, wl: Workload, steps, warmup, dist_barrier):
    """K fwd+bwd steps fed from pinned host memory with the next step's H2D copies overlapped on a side stream.
    Returns the total device milliseconds of the K steps (one CUDA-event pair around the whole region)."""
    sc, dev = wl.dev, wl.device
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    packs = [torch.empty_like(wl.host_pack, device=dev) for _ in range(2)]
    bufs = [wl.views(p) for p in packs]
    ready = [torch.cuda.Event(), torch.cuda.Event()]      # copy into buffer b finished
    free = [torch.cuda.Event(), torch.cuda.Event()]       # compute on buffer b finished
    leaf_keys = [k for k in ("means3D", "opacities", "shs", "scales", "rotations") if k in wl.host]

    def upload(b):
        with torch.cuda.stream(side):
            side.wait_event(free[b])
            packs[b].copy_(wl.host_pack, non_blocking=True)       # the step's inputs, one copy from pinned memory
            ready[b].record(side)

    marks = [0.0] * 6           # G4R_E2E_TRACE: host seconds per phase of compute()

    def compute(b):
        tr = os.environ.get("G4R_E2E_TRACE", "0") != "0"
        p0 = time.perf_counter() if tr else 0.0
        main.wait_event(ready[b])
        t = bufs[b]
        rs = dgr.GaussianRasterizationSettings(image_height=sc.H, image_width=sc.W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=t["bg"],
                                               scale_modifier=sc.scale_modifier, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"],
                                               projmatrix_raw=t["projmatrix_raw"], sh_degree=sc.sh_degree, campos=t["campos"],
                                               prefiltered=False, debug=False)
        leaf = {k: t[k].detach().requires_grad_(True) for k in leaf_keys}
        means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        p1 = time.perf_counter() if tr else 0.0
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=leaf.get("shs"),
            colors_precomp=t.get("colors_precomp"), scales=leaf.get("scales"), rotations=leaf.get("rotations"),
            cov3D_precomp=t.get("cov3D_precomp"), theta=theta, rho=rho)
        p2 = time.perf_counter() if tr else 0.0
        loss = (color * wl.grad_color).sum() + (depth * wl.grad_depth).sum()
        p3 = time.perf_counter() if tr else 0.0
        loss.backward()
        p4 = time.perf_counter() if tr else 0.0
        wl.result_host[:1].copy_(loss.detach().reshape(1), non_blocking=True)
        wl.result_host[1:4].copy_(rho.grad, non_blocking=True)
        wl.result_host[4:7].copy_(theta.grad, non_blocking=True)
        free[b].record(main)
        if tr:
            p5 = time.perf_counter()
            for k, d in enumerate((p1 - p0, p2 - p1, p3 - p2, p4 - p3, p5 - p4)):
                marks[k] += d

    for b in range(2):
        free[b].record(main)
    upload(0)
    for i in range(warmup):
        upload((i + 1) % 2)
        compute(i % 2)
    torch.cuda.synchronize()
    dist_barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace = os.environ.get("G4R_E2E_TRACE", "0") != "0"      # diagnostic: host seconds spent in upload() / compute() per step
    t_up = t_co = 0.0
    t_wall0 = time.perf_counter()
    e0.record(main)
    for i in range(warmup, warmup + steps):
        if trace:
            a = time.perf_counter()
            upload((i + 1) % 2)
            b_ = time.perf_counter()
            compute(i % 2)
            c = time.perf_counter()
            t_up += b_ - a
            t_co += c - b_
        else:
            upload((i + 1) % 2)          # the next step's inputs travel while this step computes
            compute(i % 2)
    e1.record(main)
    t_host = time.perf_counter() - t_wall0
    torch.cuda.synchronize()
    if trace:
        sys.stderr.write(f"e2e trace: host loop {1e3 * t_host / steps:.3f} ms/step (upload {1e3 * t_up / steps:.3f}, compute {1e3 * t_co / steps:.3f}); "
                         f"device {e0.elapsed_time(e1) / steps:.3f} ms/step; compute phases over warm-up + steps (ms/step): "
                         + ", ".join(f"{n} {1e3 * m / (steps + warmup):.3f}" for n, m in zip(("setup", "forward", "loss", "backward", "readback"), marks)) + "\n")
    dist_barrier()
    # informational: the same loop with the step of each buffer captured once in a CUDA graph (possible because this rasterizer
    # never blocks the host; the reference's forward cannot be captured) -- what is left is copy / kernel time, not Python
    wl.e2e_graph_ms = None
    if hasattr(dgr, "captured_overflow") and os.environ.get("G4R_E2E_NO_GRAPH", "0") == "0":
        try:
            def graph_step(b):
                t = bufs[b]
                rs = dgr.GaussianRasterizationSettings(image_height=sc.H, image_width=sc.W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=t["bg"],
                                                       scale_modifier=sc.scale_modifier, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"],
                                                       projmatrix_raw=t["projmatrix_raw"], sh_degree=sc.sh_degree, campos=t["campos"],
                                                       prefiltered=False, debug=False)
                leaf = {k: t[k].detach().requires_grad_(True) for k in leaf_keys}
                means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
                theta = torch.zeros(3, device=dev, requires_grad=True)
                rho = torch.zeros(3, device=dev, requires_grad=True)
                color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
                    means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], shs=leaf.get("shs"),
                    colors_precomp=t.get("colors_precomp"), scales=leaf.get("scales"), rotations=leaf.get("rotations"),
                    cov3D_precomp=t.get("cov3D_precomp"), theta=theta, rho=rho)
                loss = (color * wl.grad_color).sum() + (depth * wl.grad_depth).sum()
                grads = torch.autograd.grad(loss, list(leaf.values()) + [means2D, theta, rho])
                wl.result_host[:1].copy_(loss.detach().reshape(1), non_blocking=True)
                wl.result_host[1:4].copy_(grads[-1], non_blocking=True)
                wl.result_host[4:7].copy_(grads[-2], non_blocking=True)
                return grads

            cap_stream = torch.cuda.Stream(dev)
            cap_stream.wait_stream(main)
            with torch.cuda.stream(cap_stream):
                for b in range(2):
                    graph_step(b)
            main.wait_stream(cap_stream)
            torch.cuda.synchronize()
            dgr.reset_captured()
            graphs, keep = [], []
            for b in range(2):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    keep.append(graph_step(b))
                graphs.append(gph)
            torch.cuda.synchronize()
            for b in range(2):
                free[b].record(main)
            upload(0)
            for i in range(3):
                upload((i + 1) % 2)
                main.wait_event(ready[i % 2]); graphs[i % 2].replay(); free[i % 2].record(main)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(main)
            for i in range(3, 3 + steps):
                upload((i + 1) % 2)
                main.wait_event(ready[i % 2]); graphs[i % 2].replay(); free[i % 2].record(main)
            g1.record(main)
            torch.cuda.synchronize()
            if not dgr.captured_overflow():
                wl.e2e_graph_ms = g0.elapsed_time(g1) / steps
            dgr.reset_captured()
            del graphs, keep
        except Exception as exc:                     # informational leg: never take the bench line down with it
            sys.stderr.write(f"e2e cuda-graph leg skipped: {exc!r}\n")
    # the host link on its own: the same packed copy with nothing else running (tells copy-bound from kernel-bound)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        c0.record(side)
        for _ in range(5):
            packs[0].copy_(wl.host_pack, non_blocking=True)
        c1.record(side)
    torch.cuda.synchronize()
    wl.h2d_copy_ms_alone = c0.elapsed_time(c1) / 5
    return e0.elapsed_time(e1)


def run_cuda_graph(dgr, wl: Workload, steps, flush):
    """fwd + loss + bwd captured once in a CUDA graph and replayed (possible because this rasterizer never blocks the host;
    the reference's forward does a blocking cudaMemcpy and cannot be captured).  Returns per-replay milliseconds."""
    sc, dev = wl.dev, wl.device
    rs = dgr.GaussianRasterizationSettings(image_height=sc.H, image_width=sc.W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=sc.bg,
                                           scale_modifier=sc.scale_modifier, viewmatrix=sc.viewmatrix, projmatrix=sc.projmatrix,
                                           projmatrix_raw=sc.projmatrix_raw, sh_degree=sc.sh_degree, campos=sc.campos, prefiltered=False, debug=False)
    keys = [k for k in ("means3D", "opacities", "shs", "scales", "rotations") if getattr(sc, k) is not None]
    leaf = {k: getattr(sc, k).detach().clone().requires_grad_(True) for k in keys}

    def step():
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf.get("shs"), colors_precomp=sc.colors_precomp,
            scales=leaf.get("scales"), rotations=leaf.get("rotations"), cov3D_precomp=sc.cov3D_precomp, theta=theta, rho=rho)
        loss = (color * wl.grad_color).sum() + (depth * wl.grad_depth).sum()
        return torch.autograd.grad(loss, list(leaf.values()) + [m2d, theta, rho])

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    dgr.reset_captured()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3):
        graph.replay()
        flush()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        graph.replay()
        b.record()
        flush()
    torch.cuda.synchronize()
    overflow = dgr.captured_overflow()
    dgr.reset_captured()
    del out
    return [a.elapsed_time(b) for a, b in ev], overflow


def timed_steps(step, steps, warmup, flush, dist_barrier, sampler=None):
    """W warm-ups, then K steps each bracketed by CUDA events on the current stream; L2 is flushed (outside the timed
    region) between steps.  Returns per-step milliseconds."""
    if sampler:
        sampler.start()                       # nvidia-smi needs ~0.3 s to start: launch it before the warm-up
    for _ in range(warmup):
        step()
        flush()
    torch.cuda.synchronize()
    dist_barrier()
    t0 = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        step()
        b.record()
        flush()
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    dist_barrier()
    return [a.elapsed_time(b) for a, b in ev], clocks


def cpu_baseline(wl: Workload, min_s=10.0, max_frames=200, budget_s=25.0):
    """The CPU oracle (a port of the reference algorithm) on the host cores: fwd+bwd frames/s over a bounded sample
    (whole frames until at least `min_s` seconds of CPU work have been timed, never more than `budget_s`)."""
    from oracle.g4r_oracle import Oracle, scene_dict
    ora = Oracle("f32")
    d = scene_dict(wl.cpu)
    t0 = time.time()
    n = 0
    while n < max_frames and (time.time() - t0) < (min_s if n else budget_s):
        f = ora.forward(d)
        ora.backward(f, wl.cpu.grad_color, wl.cpu.grad_depth)
        n += 1
    dt = time.time() - t0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} fwd+bwd frame(s) of workload {wl.name} (P={wl.cpu.P}, {wl.cpu.W}x{wl.cpu.H}) through oracle/g4r_oracle.c (OpenMP, f32)"}



# ---------------------------------------------------------------------------------------------------------------
# the two SLAM loop shapes BASELINE.json's configs name (config 3 "inside the mapping loop"; the tracking loop of config 5)
# ---------------------------------------------------------------------------------------------------------------
def run_slam_shape(args, dgr, lib, device, rank, world, barrier, ref_cuda):
    """--workload C3map : one step = one iteration of BackEnd.map (utils/slam_backend.py:357-771): 10 views of the 500 k cloud
                          rendered through render(), RGB-D mapping losses, ONE backward(retain_graph=True) through the 10
                          rasterizer nodes, Adam step on the Gaussians + the 10 poses.  metric = views/s (10 per step).
       --workload track : one step = one iteration of FrontEnd.tracking (utils/slam_frontend.py:411-448): render() of a
                          60 k-Gaussian map with the static mask (30 k kept), RGB-D tracking loss, backward, pose Adam step,
                          SE(3) update; only theta / rho are consumed.
    Both through tools/slam_shapes (a restatement of the reference caller; tests/test_reference_caller.py runs the reference's
    own file on the same objects).  e2e: the ground-truth colour + depth images of the step's views come from pinned host
    memory every step and the loss goes back."""
    from tools import slam_shapes as S
    from tools.scenes import config_scene, make_scene
    mapping = args.workload == "C3map"
    views_n = 10 if mapping else 1
    sc_cpu = config_scene("C3", seed=rank) if mapping else make_scene(60_000, 640, 480, sh_degree=0, seed=rank, name="track60k")
    sc = sc_cpu.to(device)
    g = torch.Generator().manual_seed(5 + rank)
    dygs = torch.zeros(sc.P, dtype=torch.bool) if mapping else (torch.rand(sc.P, generator=g) < 0.5)
    gt_host = [(torch.rand(3, sc.H, sc.W, generator=g).pin_memory(), (0.5 + 5.0 * torch.rand(1, sc.H, sc.W, generator=g)).pin_memory())
               for _ in range(views_n)]
    bg = torch.zeros(3, device=device)
    pc = S.DuckGaussians.from_scene(sc, dygs=dygs.to(device))
    base = S.DuckCamera.from_scene(sc)
    views = [S.perturbed(base, 10 + i, rot=0.004, trans=0.01) for i in range(views_n)]
    for v, (im, dp) in zip(views, gt_host):
        v.original_image, v.depth_gt = im.to(device), dp.to(device)
    render_fn = lambda *a, **k: S.render(dgr, *a, **k)
    gopt = pc.optimizer()
    popts = [S.pose_optimizer(v) for v in views]
    result_host = torch.empty(1, dtype=torch.float32).pin_memory()

    fused = [False]
    frozen = tuple(p.detach() for p in pc.parameters())           # tracking never updates the map
    static_mask = (pc.dygs == False).to(torch.uint8)              # noqa: E712

    def step(from_host=False):
        if from_host:
            for v, (im, dp) in zip(views, gt_host):
                v.original_image, v.depth_gt = im.to(device, non_blocking=True), dp.to(device, non_blocking=True)
        if fused[0]:
            loss = (S.mapping_iteration_fused(dgr, views, pc, bg, gopt, popts) if mapping
                    else S.tracking_iteration_fused(dgr, views[0], frozen, static_mask, bg, popts[0]))
        elif mapping:
            loss = S.mapping_iteration(dgr, render_fn, views, pc, bg, gopt, popts)
        else:
            loss = S.tracking_iteration(dgr, render_fn, views[0], pc, bg, popts[0], gopt)
        if from_host:
            result_host.copy_(loss.detach().reshape(1), non_blocking=True)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flush = lambda: flush_buf.zero_()
    steps = args.steps if not mapping else max(5, args.steps // 4)
    if lib is not None:
        lib.g4r_profile_enable(1)
        n_st = lib.g4r_profile_stage_count()
        lib.g4r_profile_stage_name.restype = ctypes.c_char_p
    ms, clocks = timed_steps(step, steps, args.warmup, flush, barrier, ClockSampler(torch.cuda.current_device()) if rank == 0 else None)
    stage = {}
    if lib is not None:
        ms_arr = (ctypes.c_double * n_st)()
        cnt_arr = (ctypes.c_int64 * n_st)()
        lib.g4r_profile_read(ms_arr, cnt_arr, 1)
        lib.g4r_profile_enable(0)
        stage = {lib.g4r_profile_stage_name(i).decode(): (ms_arr[i] / max(1, cnt_arr[i]), int(cnt_arr[i])) for i in range(n_st)}
    t_total = torch.tensor([sum(ms)], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms = float(t_total.item())
    value = world * steps * views_n / (total_ms / 1000.0)
    # e2e: host ground-truth images every step, loss read back; one event pair around K steps
    K2 = max(5, steps // 2)
    for _ in range(3):
        step(True)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K2):
        step(True)
    e1.record()
    torch.cuda.synchronize()
    t_h = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_h, op=dist.ReduceOp.MAX)
    e2e_value = world * K2 * views_n / (float(t_h.item()) / 1000.0)
    # the same loop body on the opt-in fused API of this repo (raw parameters + in-kernel mask, fused loss, pose-only backward)
    fused_rec = None
    if lib is not None:
        fused[0] = True
        fms, _ = timed_steps(step, steps, args.warmup, flush, barrier, None)
        t_f = torch.tensor([sum(fms)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
        fused_rec = {"value": world * steps * views_n / (float(t_f.item()) / 1000.0), "unit": "frames/s", "ms_per_step": float(t_f.item()) / steps,
                     "api": "FusedGaussianRasterizer (raw parameters, in-kernel static mask) + slam_loss (one kernel) + "
                            + ("pose-only backward" if not mapping else "Adam") + "; camera matrices evaluated once per render"}
        fused[0] = False
    if rank != 0:
        return
    h2d = sum(im.numel() * 4 + dp.numel() * 4 for im, dp in gt_host)
    peak, peak_src = measured_peaks()
    ms_per_step = total_ms / steps
    visible = int((~dygs).sum())
    line = {
        "metric": ("mapping-loop rasterizer views/s: 10 views fwd, one backward(retain_graph=True), Adam; 500k Gaussians @640x480" if mapping
                   else "tracking-loop iterations/s: masked render (30k of 60k Gaussians) fwd+bwd, pose Adam step @640x480"),
        "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: P={sc.P} Gaussians ({visible} after the static mask), {sc.W}x{sc.H}, SH degree 0, {views_n} view(s) per step, "
                               "render() prelude + RGB-D loss + backward + Adam through tools/slam_shapes (restating utils/slam_backend.py:357-771 / "
                               "utils/slam_frontend.py:411-448)", "parallelism": f"replicas x{world}",
                   "l2": "flushed between steps (256 MiB memset outside the timed events)",
                   "api": ("UNMODIFIED reference build baseline/_ref" if ref_cuda else "this repo's drop-in") + " behind the reference's render() call shape"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(sum(c for _, c in stage.values())) if stage else 0,
        "clocks": clocks,
    }
    if ref_cuda:
        line["impl"] = "reference"
    if fused_rec is not None:
        line["fused_api"] = fused_rec
    if stage:
        line["kernel_ms"] = {k: round(v[0], 5) for k, v in stage.items()}
        line["kernel_launches_per_step"] = {k: round(v[1] / (steps + args.warmup), 2) for k, v in stage.items()}
        native = sum(v[0] * v[1] for v in stage.values()) / (steps + args.warmup)
        line["native_kernel_ms_per_step"] = native
        line["non_rasterizer_ms_per_step"] = ms_per_step - native
    print(json.dumps(line))

# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path (the CPU oracle is only the baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        barrier = lambda: dist.barrier(device_ids=[local_rank])
    else:
        barrier = lambda: None

    from tools import refload
    ref_cuda = args.impl == "reference" and refload.available()
    if args.impl == "reference" and not ref_cuda:
        # no reference build on this box: the reference arm is the CPU port of its algorithm, rank 0 only
        if rank == 0:
            wl = Workload(args.workload, seed=0, device=device)
            cb = cpu_baseline(wl, max_frames=max(1, min(args.steps, 3)))
            print(json.dumps({"impl": "reference", "metric": "rasterizer fwd+bwd frames/sec", "value": cb["value"], "unit": "frames/s",
                              "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / cb["value"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": args.workload, "note": "baseline/_ref absent: CPU oracle port timed instead"},
                              "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        if use_dist:
            barrier()
        return

    if ref_cuda:
        dgr = refload.load()
        lib = None
    else:
        import diff_gaussian_rasterization as dgr
        lib = dgr._lib
    if args.workload in ("C3map", "track"):
        run_slam_shape(args, dgr, lib, device, rank, world, barrier, ref_cuda)
        if use_dist:
            barrier()
            dist.destroy_process_group()
        return
    # every rank renders its own view (seed) of the workload
    wl = Workload(args.workload, seed=rank, device=device)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)       # > 126 MB L2
    flush = lambda: flush_buf.zero_()

    # ---- device-resident leg (value) ------------------------------------------------------------------------
    step = make_step(dgr, wl, from_host=False)
    if lib is not None:
        lib.g4r_profile_enable(1)
        n_st = lib.g4r_profile_stage_count()
        lib.g4r_profile_stage_name.restype = ctypes.c_char_p
    ms, clocks = timed_steps(step, args.steps, args.warmup, flush, barrier, ClockSampler(local_rank) if rank == 0 else None)
    stage = {}
    if lib is not None:
        ms_arr = (ctypes.c_double * n_st)()
        cnt_arr = (ctypes.c_int64 * n_st)()
        lib.g4r_profile_read(ms_arr, cnt_arr, 1)
        lib.g4r_profile_enable(0)
        # the profile covers warm-up + timed steps alike (same work every step)
        stage = {lib.g4r_profile_stage_name(i).decode(): (ms_arr[i] / max(1, cnt_arr[i]), int(cnt_arr[i])) for i in range(n_st)}
    t_total = torch.tensor([sum(ms)], dtype=torch.float64, device=device)
    if use_dist:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms = float(t_total.item())
    value = world * args.steps / (total_ms / 1000.0)

    # ---- end-to-end leg (host buffers) ------------------------------------------------------------------------
    # Every step's inputs come from pinned host memory and its result (loss + pose gradient) goes back to the host.  The
    # copies of step i+1 are issued on a side stream while step i computes (double-buffered device inputs), so all K
    # H2D/D2H transfers happen inside the single timed region of K steps.  No L2 flush here: each step streams ~28 MB of
    # fresh inputs plus ~150 MB of scratch through the 126 MB L2.
    K2 = max(5, args.steps // 2)
    e2e_ms = run_e2e(dgr, wl, K2, 3, barrier)
    t_h = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
    if use_dist:
        dist.all_reduce(t_h, op=dist.ReduceOp.MAX)
    e2e_value = world * K2 / (float(t_h.item()) / 1000.0)

    # ---- Gaussian-sharded render of ONE large frame (BASELINE.json config 4; N > 1 only) -----------------------------------
    # The headline `value` above is the C3 frame replicated over independent views (no data-path collective).  The path
    # north_star shards is the 2 M-Gaussian 1280x960 mapping render: every rank owns a shard of the Gaussians and a strip of
    # tile rows; NCCL all-to-all of splat records, all-gather of image strips, reverse all-to-all of gradient rows.
    sharded_rec = None
    if use_dist and lib is not None and args.workload == "C3":
        try:
            from tools.sharded_check import measure, passed
            rep = measure("X4", 10, device, rank, world, local_rank)
            reps = [None] * world
            dist.all_gather_object(reps, rep)
            sharded_rec = dict(reps[0])
            sharded_rec["parity_on_every_rank"] = all(passed(r) for r in reps)
            sharded_rec["note"] = ("X4 = 2 M Gaussians, 1280x960, SH degree 0 on the bit-reproducible generator (tools.scenes.exact_scene); "
                                   "ms = fwd+bwd through ShardedGaussianRasterizer, CUDA events, max over ranks; single-GPU ms = the same "
                                   "frame through GaussianRasterizer on one GPU of this box")
        except Exception as exc:                      # the replica numbers stand on their own
            sys.stderr.write(f"sharded leg failed: {exc!r}\n")
            sharded_rec = {"error": repr(exc)[:300]}

    graph_ms, graph_overflow = (None, None)
    if lib is not None and rank == 0 and not use_dist:
        try:
            gms, graph_overflow = run_cuda_graph(dgr, wl, args.steps, flush)
            graph_ms = sum(gms) / len(gms)
        except Exception as exc:                      # the eager numbers stand on their own
            sys.stderr.write(f"cuda-graph leg skipped: {exc!r}\n")

    if rank == 0:
        sc = wl.cpu
        # workload statistics for the roofline: N from one extra forward
        if lib is not None:
            from tools import runners
            _, info = dgr.rasterize_gaussians_with_state(runners.settings_for(wl.dev, dgr), wl.dev.means3D, wl.dev.opacities, shs=wl.dev.shs,
                                                         colors_precomp=wl.dev.colors_precomp, scales=wl.dev.scales, rotations=wl.dev.rotations,
                                                         cov3D_precomp=wl.dev.cov3D_precomp)
            N = int(info["num_rendered"])
        else:
            N = int(refload.run_reference(wl.dev, want_grads=False)["num_rendered"])
        K = (sc.sh_degree + 1) ** 2
        X = sc.W * sc.H
        tiles = ((sc.W + 15) // 16) * ((sc.H + 15) // 16)
        b_alg, per_kernel, S = algorithmic_bytes(sc.P, K, N, X, tiles)
        peak, peak_src = measured_peaks()
        ms_per_step = total_ms / args.steps
        line = {
            "metric": "rasterizer fwd+bwd frames/sec @640x480, 500k Gaussians; HBM GB/s vs peak" if args.workload == "C3"
                      else f"rasterizer fwd+bwd frames/sec, workload {args.workload}",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: P={sc.P} Gaussians, {sc.W}x{sc.H}, SH degree {sc.sh_degree} (M={K}), fwd+bwd incl. pose grads, "
                                   f"num_rendered N={N}", "parallelism": f"views x{world} (independent frames per GPU, no data-path collective)",
                       "l2": "flushed between steps (256 MiB memset outside the timed events)", "timing": "CUDA events per step on the current stream, "
                       "sum over K steps, max over ranks",
                       "api": "public GaussianRasterizer autograd API (host side: " + (dgr.host_backend() if hasattr(dgr, "host_backend") else "reference") +
                              " -> C ABI)" if lib is not None else "public GaussianRasterizer autograd API of the reference build"},
            # e2e streams the frame's 28 MB of Gaussian tensors from pinned host memory every step (one packed copy,
            # double-buffered on a side stream).  h2d_link_gbs_alone = the same copy with nothing else running: at C3 the link
            # needs 0.51 ms per step, the kernels 0.62 ms, and the Python / autograd host work of one eager step ~0.78 ms -- the
            # e2e loop of this arm is host-bound (the reference arm is kernel-bound at 5.3 ms); N ranks also share the host's
            # DRAM / root complexes.
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes,
                    "h2d_gbs_per_gpu": e2e_value / world * wl.h2d_bytes / 1e9, "h2d_gbs_aggregate": e2e_value * wl.h2d_bytes / 1e9,
                    "cuda_graph_value": (world * 1000.0 / wl.e2e_graph_ms) if getattr(wl, "e2e_graph_ms", None) else None,
                    "cuda_graph_note": "informational: the same double-buffered loop with each buffer's step captured in a CUDA graph "
                                       "(this rank's rate x ranks); `value` above is the eager loop",
                    "h2d_copy_ms_alone": getattr(wl, "h2d_copy_ms_alone", None),
                    "h2d_link_gbs_alone": (wl.h2d_bytes / 1e6 / wl.h2d_copy_ms_alone) if getattr(wl, "h2d_copy_ms_alone", None) else None},
            # launches of this library's kernels inside the timed region, counted by the library's stage profile (it covers
            # the warm-up steps too, which run the same launches)
            "gpu_launches": int(round(sum(v[1] for v in stage.values()) * args.steps / (args.steps + args.warmup))) if stage else 0,
            "clocks": clocks,
            "roofline_frame": {"bound": "hbm", "algorithmic_bytes": b_alg, "achieved": b_alg / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                               "unit": "GB/s", "frac": b_alg / (ms_per_step * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                               "sort_passes_S": S},
        }
        if ref_cuda:
            line["impl"] = "reference"
            line["config"]["api"] = "UNMODIFIED reference build baseline/_ref (sm_100a) through its own GaussianRasterizer API"
        if stage:
            groups = {"project": ("project",), "binning": ("tile_scan", "scatter", "tile_sort"), "composite_forward": ("composite_forward",),
                      "composite_backward": ("composite_backward",), "gaussian_backward": ("gaussian_backward",)}
            kt = {g: sum(stage[s][0] for s in members) for g, members in groups.items()}
            dom = max(kt, key=kt.get)
            ach = per_kernel[dom] / (kt[dom] * 1e-3) / 1e9
            traffic = None
            warp_inst = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath) and args.workload == "C3":
                # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed `ncu --set full` capture
                with open(tpath) as f:
                    tj = json.load(f)
                for kname, kv in tj["kernels"].items():
                    if dom in kname or (dom == "binning" and "tile_sort" in kname):
                        traffic = kv["dram_bytes_per_launch"]
                        warp_inst = kv.get("warp_instructions")
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": traffic, "algorithmic_bytes_per_launch": per_kernel[dom], "ms_per_launch": kt[dom],
                                "peak_source": peak_src}
            line["kernel_ms"] = {k: round(v[0], 5) for k, v in stage.items()}
            if warp_inst and clocks and clocks.get("sm_mhz"):
                # The dominant kernel is bound by instruction issue, not by HBM: warp instructions it executes
                # (smsp__inst_executed.sum of the committed ncu capture) per second of its live duration, against
                # SMs x 4 schedulers x the SM clock sampled during this run.  Informational, next to the HBM roofline above.
                n_sm = torch.cuda.get_device_properties(device).multi_processor_count
                peak_issue = n_sm * 4 * clocks["sm_mhz"] * 1e6
                ach_issue = warp_inst / (kt[dom] * 1e-3)
                line["issue_roofline"] = {"kernel": dom, "warp_instructions_per_launch": warp_inst, "achieved": ach_issue / 1e9,
                                          "peak": peak_issue / 1e9, "unit": "G warp-instructions/s", "frac": ach_issue / peak_issue}
        if sharded_rec is not None:
            line["sharded"] = sharded_rec
        if graph_ms is not None:
            line["cuda_graph"] = {"value": 1000.0 / graph_ms, "unit": "frames/s", "ms_per_step": graph_ms, "capacity_overflow": graph_overflow,
                                  "note": "same step (fwd + loss + bwd through the public API) captured once with torch.cuda.graph and replayed; "
                                          "informational -- `value` above is the eager number"}
        if not args.no_cpu_baseline and world == 1:          # rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
    if use_dist:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
