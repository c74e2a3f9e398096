"""Run the UNMODIFIED reference simple-knn build (baseline/_ref/simple_knn, installed by __graft_entry__.build_reference_knn) on
the bit-reproducible clouds of tools/knn_cases.py and record sha256 digests of its output; time both implementations.

    python tools/knn_digests.py --write tests/golden/knn_digests_ref.json       # on a GPU box that has baseline/_ref
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "4dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from tools import knn_cases  # noqa: E402


def load_reference():
    """baseline/_ref/simple_knn/_C*.so, loaded under a private name so that it cannot shadow (or be shadowed by) ours."""
    import glob
    import importlib.util
    so = glob.glob(os.path.join(ROOT, "baseline", "_ref", "simple_knn", "_C*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("_C", so[0])      # the module's init symbol is PyInit__C
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def time_ms(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "knn_bench.json"))
    ap.add_argument("names", nargs="*", default=list(knn_cases.EXACT_CLOUDS))
    args = ap.parse_args()
    ref = load_reference()
    from simple_knn._C import distCUDA2
    dev = torch.device("cuda", 0)
    digests, bench = {}, {}
    for name in args.names:
        pts_np = knn_cases.exact_cloud(name)
        pts = torch.from_numpy(pts_np).to(dev)
        mine = distCUDA2(pts)
        row = dict(P=int(pts.shape[0]), ms_ours=time_ms(lambda: distCUDA2(pts)))
        if ref is not None:
            theirs = ref.distCUDA2(pts)
            row["bit_identical_to_reference"] = bool(torch.equal(mine.view(torch.int32), theirs.view(torch.int32)))
            row["ms_reference"] = time_ms(lambda: ref.distCUDA2(pts))
            row["speedup"] = row["ms_reference"] / row["ms_ours"]
            a = theirs.cpu().numpy()
            digests[name] = dict(P=int(a.size), input_sha256=knn_cases.sha(pts_np), sha256=knn_cases.sha(a), sum=float(a.astype(np.float64)[np.isfinite(a)].sum()))
        bench[name] = row
        print(name, row, flush=True)
    # where the time goes on the largest cloud (kernel table of one call, both implementations)
    from torch.profiler import ProfilerActivity, profile
    big = torch.from_numpy(knn_cases.exact_cloud(args.names[-1])).to(dev)
    for tag, fn in (("ours", distCUDA2),) + ((("reference", ref.distCUDA2),) if ref is not None else ()):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(big)
            torch.cuda.synchronize()
        with open(os.path.join(os.path.dirname(args.out), f"knn_kernels_{tag}.txt"), "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(bench, open(args.out, "w"), indent=1)
    if args.write and ref is not None:
        json.dump(digests, open(args.write, "w"), indent=1)
        print("wrote", args.write)


if __name__ == "__main__":
    main()
