"""Point clouds for the distCUDA2 (simple-knn) parity checks.

exact_cloud(name): bit-reproducible on every machine (SplitMix64 integers + float64 + - * only, then one cast to float32), so
that sha256 digests of the UNMODIFIED reference build's output recorded on a B200 (tests/golden/knn_digests_ref.json, written by
tools/knn_digests.py) can be compared anywhere.  random_cases(): shapes that stress the search (outliers, planes, lines,
duplicates, lattices, tiny P), checked against the brute-force oracle."""
from __future__ import annotations

import ctypes
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EXACT_CLOUDS = {           # name -> (kind, P)
    "K1_depthmap_20k": ("depthmap", 20_000),
    "K2_uniform_100k": ("uniform", 100_000),
    "K3_depthmap_500k": ("depthmap", 500_000),
    "K4_outliers_200k": ("outliers", 200_000),
    "K5_duplicates_60k": ("duplicates", 60_000),
    "K6_depthmap_2M": ("depthmap", 2_000_000),
}


def exact_cloud(name: str) -> np.ndarray:
    from tools.scenes import _splitmix_uniform
    kind, P = EXACT_CLOUDS[name]
    seed = 4000 + list(EXACT_CLOUDS).index(name)
    U = lambda k, n=P: _splitmix_uniform(seed, k, n)
    if kind == "uniform":
        pts = np.stack([U(1), U(2), U(3)], 1) * 4.0 - 2.0
    elif kind == "depthmap":
        # what the reference seeds from (gaussian_model.py:186-240): pixels of an RGB-D frame back-projected through a pinhole
        # camera -- a 2-D surface in 3-D with density falling off with depth, plus a little noise
        u, v = U(1) * 2.0 - 1.0, U(2) * 1.5 - 0.75
        z = 2.0 + u * v + 0.5 * u * u + (v * v) * (v + 1.0) + 0.01 * U(3)
        z = np.where(U(4) < 0.2, z + 3.0, z)                          # a second surface further away
        pts = np.stack([u * z, v * z, z], 1)
    elif kind == "outliers":
        pts = np.stack([U(1), U(2), U(3)], 1) * 0.5
        n_out = P // 1000
        pts[:n_out] = (np.stack([U(5, n_out), U(6, n_out), U(7, n_out)], 1) - 0.5) * 2000.0
    elif kind == "duplicates":
        base = np.stack([U(1, P // 3), U(2, P // 3), U(3, P // 3)], 1)
        pts = np.concatenate([base, base, base[: P - 2 * (P // 3)]], 0)
    else:
        raise KeyError(kind)
    return np.ascontiguousarray(pts.astype(np.float32))


def random_cases(seed: int = 0, n: int = 20000):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    yield "uniform", rng.random((n, 3), dtype=f32)
    yield "gaussian", rng.standard_normal((n, 3)).astype(f32)
    s = rng.standard_normal((n, 3)).astype(f32)
    yield "sphere", (s / np.linalg.norm(s, axis=1, keepdims=True)).astype(f32)
    a = rng.standard_normal((n, 3)).astype(f32) * f32(0.01)
    a[:20] = rng.standard_normal((20, 3)).astype(f32) * f32(1000)
    yield "far_outliers", a
    p = rng.random((n, 3), dtype=f32); p[:, 2] = 0.5
    yield "plane", p
    l = rng.random((n, 3), dtype=f32); l[:, 1:] = 0
    yield "line", l
    d = rng.random((n // 4, 3), dtype=f32)
    yield "duplicates", np.concatenate([d, d, d, d[:100]])
    yield "all_identical", np.ones((n // 4, 3), f32)
    g = np.stack(np.meshgrid(*[np.arange(27, dtype=f32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    yield "lattice_ties", g
    yield "offset_1000", (rng.random((n, 3), dtype=f32) + f32(1000)).astype(f32)
    yield "lognormal", np.exp(rng.standard_normal((n, 3)) * 2).astype(f32)
    yield "negative", (rng.random((n, 3), dtype=f32) * f32(-5)).astype(f32)
    for P in (1, 2, 3, 4, 5, 7, 9, 33, 257):
        yield f"tiny_{P}", rng.random((P, 3), dtype=f32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def host_search(points: np.ndarray, per_query: bool = False):
    """The product's search code (csrc/knn_search.cuh) compiled for the HOST (tests/native/knn_search_host.cpp): test harness."""
    from oracle import g4r_oracle
    g4r_oracle.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libknn_search_host.so"))
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    out = np.empty(pts.shape[0], np.float32)
    stats = np.zeros(4, np.uint64)
    evals = np.zeros((pts.shape[0] if per_query else 0, 2), np.uint32)      # per query: distance evaluations, nodes visited
    lib.knn_host_mean_dist2(ctypes.c_int32(pts.shape[0]), ctypes.c_void_p(pts.ctypes.data), ctypes.c_void_p(out.ctypes.data),
                            ctypes.c_void_p(stats.ctypes.data), ctypes.c_void_p(evals.ctypes.data if per_query else None))
    return out, dict(evals=evals if per_query else None, evals_per_query=float(stats[0]) / max(1, pts.shape[0]), cells_per_query=float(stats[1]) / max(1, pts.shape[0]),
                     rounds_per_query=float(stats[2]) / max(1, pts.shape[0]), nodes_per_query=float(stats[3]) / max(1, pts.shape[0]))
