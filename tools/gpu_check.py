"""GPU differential check (run under gpurun): this repo's kernels vs the unmodified reference build (baseline/_ref)
vs the CPU oracle on seeded scenes; optionally dumps the reference's outputs as golden fixtures.

    python tools/gpu_check.py [--golden] [--big] [--out gpurun_out]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import refload, runners  # noqa: E402
from tools.scenes import make_scene, config_scene  # noqa: E402

GOLDEN_SCENES = {
    # name: kwargs of make_scene
    "g1_deg0": dict(P=1200, W=128, H=96, sh_degree=0, seed=11),
    "g2_deg3": dict(P=1200, W=128, H=96, sh_degree=3, seed=12),
    "g3_precomp": dict(P=1000, W=100, H=75, sh_degree=0, seed=13, colors_precomp=True, cov3D_precomp=True),
    "g4_deg1_m16": dict(P=1200, W=128, H=96, sh_degree=1, sh_coeffs=16, seed=14, scale_modifier=1.3),
    "g5_big_splats": dict(P=600, W=128, H=96, sh_degree=0, seed=15, px_min=3.0, px_max=20.0),
}


def to_np(d: dict) -> dict:
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
        elif isinstance(v, (int, float, np.ndarray)):
            out[k] = np.asarray(v)
    return out


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    dev = torch.device("cuda:0")
    print("device:", torch.cuda.get_device_name(0), "reference available:", refload.available(), flush=True)
    report = {}

    scenes = {k: make_scene(name=k, **kw) for k, kw in GOLDEN_SCENES.items()}
    if args.big:
        scenes["C2"] = config_scene("C2")
        scenes["C3"] = config_scene("C3")
        scenes["C3_identity_cam"] = make_scene(500_000, 640, 480, sh_degree=0, seed=5, posed=False, off_centre=False, name="C3_identity_cam")

    for name, sc_cpu in scenes.items():
        print(f"=== {name}: P={sc_cpu.P} {sc_cpu.W}x{sc_cpu.H} deg={sc_cpu.sh_degree}", flush=True)
        rep = {}
        try:
            sc = sc_cpu.to(dev)
            t0 = time.time()
            mine = runners.run_g4r(sc)
            torch.cuda.synchronize()
            print(f"  ours ran in {time.time() - t0:.2f}s  N={mine['num_rendered']}", flush=True)
            if refload.available():
                ref = refload.run_reference(sc)
                torch.cuda.synchronize()
                rep["ours_vs_reference"] = runners.compare(mine, ref)
                print("  ours vs reference:\n" + runners.fmt_report(rep["ours_vs_reference"]), flush=True)
                # the reference against itself (atomics order noise floor for the gradients)
                ref2 = refload.run_reference(sc)
                rep["reference_vs_reference"] = {k: v for k, v in runners.compare(ref2, ref).items() if k.startswith("dL_")}
                if args.golden and name in GOLDEN_SCENES:
                    gd = os.path.join(args.out, "golden")
                    os.makedirs(gd, exist_ok=True)
                    keep = ("color", "depth", "opacity", "radii", "n_touched", "num_rendered", "point_list", "ranges", "n_contrib",
                            "final_T", "means2D", "depths", "conic_opacity", "rgb", "clamped", "dL_dmeans3D", "dL_dmeans2D",
                            "dL_dopacity", "dL_dshs", "dL_dcolors", "dL_dscales", "dL_drots", "dL_dcov3D", "dL_dtau")
                    blob = {"ref_" + k: v for k, v in to_np({k: ref[k] for k in keep if k in ref}).items()}
                    for k, v in sc_cpu.__dict__.items():
                        if isinstance(v, torch.Tensor):
                            blob["in_" + k] = v.numpy()
                    blob["in_scalars"] = np.array([sc_cpu.W, sc_cpu.H, sc_cpu.sh_degree, sc_cpu.tanfovx, sc_cpu.tanfovy, sc_cpu.scale_modifier],
                                                  dtype=np.float64)
                    np.savez_compressed(os.path.join(gd, name + ".npz"), **blob)
            if sc_cpu.P <= 200_000:
                t0 = time.time()
                ora = runners.run_oracle(sc_cpu)
                print(f"  oracle ran in {time.time() - t0:.2f}s", flush=True)
                rep["ours_vs_oracle"] = runners.compare(mine, ora)
                print("  ours vs oracle:\n" + runners.fmt_report(rep["ours_vs_oracle"]), flush=True)
                if refload.available():
                    rep["oracle_vs_reference"] = runners.compare(ora, ref)
                    print("  oracle vs reference:\n" + runners.fmt_report(rep["oracle_vs_reference"]), flush=True)
            if args.big and name.startswith("C"):
                import diff_gaussian_rasterization as dgr
                t_mine = timeit(lambda: runners.run_public_api(sc, dgr))
                rep["ms_fwd_bwd_ours"] = t_mine
                if refload.available():
                    t_ref = timeit(lambda: runners.run_public_api(sc, refload.load()))
                    rep["ms_fwd_bwd_reference"] = t_ref
                print(f"  fwd+bwd ms: ours {t_mine:.3f}  reference {rep.get('ms_fwd_bwd_reference', float('nan')):.3f}", flush=True)
        except Exception:
            rep["error"] = traceback.format_exc()
            print(rep["error"], flush=True)
        report[name] = rep
        with open(os.path.join(args.out, "gpu_check.json"), "w") as f:
            json.dump(report, f, indent=1, default=str)
    print("done")


if __name__ == "__main__":
    main()
