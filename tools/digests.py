"""Digests of rasterizer outputs: what tests/golden/digests_ref.json stores for the UNMODIFIED reference build at BASELINE.json's
full sizes, so that full-size parity does not depend on the git-ignored baseline/_ref travelling to the test box.

    python tools/digests.py --write tests/golden/digests_ref.json [X1 X2 X3 X4]     # on a GPU box that has baseline/_ref

Scenes come from tools.scenes.exact_scene (bit-reproducible inputs; their sha256 is stored and re-checked by the tests).
Per scene: sha256 of every integer output and of the bits of every image; per gradient tensor its L2 norm, its sum and its
projections onto four fixed +-1 vectors (SplitMix64 signs) -- a gradient that differs from the reference's by a relative L2
error e moves a projection by about e * norm, so the tests bound |projection difference| by 1e-3 * norm."""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "4dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

INT_KEYS = ("radii", "n_touched", "point_list", "ranges", "n_contrib")
IMG_KEYS = ("color", "depth", "opacity", "final_T")
GRAD_KEYS = ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dshs", "dL_dscales", "dL_drots", "dL_dtau")
N_PROJ = 4


def _np(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def _signs(n: int, k: int) -> np.ndarray:
    from tools.scenes import _splitmix_uniform
    return np.where(_splitmix_uniform(977, 100 + k, n) < 0.5, -1.0, 1.0)


def digest(out: dict) -> dict:
    d = {"num_rendered": int(out["num_rendered"])}
    for k in INT_KEYS:
        a = np.ascontiguousarray(_np(out[k]).astype(np.int32).reshape(-1))
        d[k] = {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "size": int(a.size), "sum": int(a.astype(np.int64).sum())}
    for k in IMG_KEYS:
        a = np.ascontiguousarray(_np(out[k]).astype(np.float32).reshape(-1))
        d[k] = {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "size": int(a.size), "sum": float(a.astype(np.float64).sum())}
    for k in GRAD_KEYS:
        if out.get(k) is None:
            continue
        a = _np(out[k]).astype(np.float64).reshape(-1)
        d[k] = {"size": int(a.size), "norm": float(np.linalg.norm(a)), "sum": float(a.sum()),
                "proj": [float(a @ _signs(a.size, j)) for j in range(N_PROJ)]}
    return d


def compare(mine: dict, ref: dict, grad_tol: float = 1e-3) -> list:
    """Differences of digest `mine` from the reference digest `ref` (empty list = parity)."""
    bad = []
    if mine["num_rendered"] != ref["num_rendered"]:
        bad.append(("num_rendered", mine["num_rendered"], ref["num_rendered"]))
    for k in INT_KEYS + IMG_KEYS:
        if mine[k]["sha256"] != ref[k]["sha256"]:
            bad.append((k, mine[k]["sum"], ref[k]["sum"]))
    for k in GRAD_KEYS:
        if k not in ref:
            continue
        n = ref[k]["norm"]
        if abs(mine[k]["norm"] - n) > grad_tol * n:
            bad.append((k + ".norm", mine[k]["norm"], n))
        for a, b in zip(mine[k]["proj"], ref[k]["proj"]):
            if abs(a - b) > grad_tol * n:
                bad.append((k + ".proj", a, b))
    return bad


def main():
    from tools import refload
    from tools.scenes import EXACT_CONFIGS, exact_scene, input_digest
    path = sys.argv[sys.argv.index("--write") + 1]
    names = [a for a in sys.argv[1:] if a in EXACT_CONFIGS] or list(EXACT_CONFIGS)
    dev = torch.device("cuda:0")
    res = json.load(open(path)) if os.path.exists(path) else {}
    res["_about"] = ("outputs of the UNMODIFIED reference build (baseline/_ref, sm_100a) on tools.scenes.exact_scene inputs; "
                     "written by tools/digests.py on a B200")
    for name in names:
        sc_cpu = exact_scene(name)
        sc = sc_cpu.to(dev)
        ref = refload.run_reference(sc)
        ref2 = refload.run_reference(sc)           # the reference's own run-to-run gradient noise (float atomics order)
        d = digest(ref)
        d["input_sha256"] = input_digest(sc_cpu)
        d["reference_self_noise"] = {k: float((ref[k].double() - ref2[k].double()).norm() / (ref[k].double().norm() + 1e-30))
                                     for k in GRAD_KEYS if ref.get(k) is not None}
        res[name] = d
        print(name, "N =", d["num_rendered"], "input", d["input_sha256"][:12], flush=True)
        del ref, ref2, sc
        torch.cuda.empty_cache()
    with open(path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
