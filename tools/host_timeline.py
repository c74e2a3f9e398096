"""Where does the eager (non-CUDA-graph) time go on small scenes?  Host-side timeline of the public API call sequence
(run under gpurun, 1 GPU): wall time per iteration, host time inside forward / loss / backward, host time inside every
native entry point, number of phase-2 launches per frame (a second one means the speculative capacity overflowed)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import refload, runners  # noqa: E402
from tools.scenes import config_scene, make_scene  # noqa: E402


class Timed:
    def __init__(self, fn, name, acc):
        self.fn, self.name, self.acc = fn, name, acc

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        d = self.acc.setdefault(self.name, [0.0, 0])
        d[0] += time.perf_counter() - t0
        d[1] += 1
        return r


def loop(sc, dgr, iters, fresh_leaves=True):
    """Returns (wall ms / iteration, host ms in forward, loss, backward)."""
    rs = runners.settings_for(sc, dgr)
    dev = sc.means3D.device
    keys = [k for k in ("means3D", "opacities", "shs", "scales", "rotations") if getattr(sc, k) is not None]
    leaf = {k: getattr(sc, k).detach().clone().requires_grad_(True) for k in keys}
    h = [0.0, 0.0, 0.0, 0.0]
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    for _ in range(iters):
        t0 = time.perf_counter()
        if fresh_leaves:
            leaf = {k: getattr(sc, k).detach().clone().requires_grad_(True) for k in keys}
        else:
            for v in leaf.values():
                v.grad = None
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        t1 = time.perf_counter()
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf.get("shs"), colors_precomp=sc.colors_precomp,
            scales=leaf.get("scales"), rotations=leaf.get("rotations"), cov3D_precomp=sc.cov3D_precomp, theta=theta, rho=rho)
        t2 = time.perf_counter()
        loss = (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()
        t3 = time.perf_counter()
        loss.backward()
        t4 = time.perf_counter()
        h[0] += t1 - t0; h[1] += t2 - t1; h[2] += t3 - t2; h[3] += t4 - t3
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t_begin) / iters * 1e3
    return wall, [x / iters * 1e3 for x in h]


def main():
    dev = torch.device("cuda:0")
    import diff_gaussian_rasterization as dgr
    ref = refload.load() if refload.available() else None
    cases = {
        "C2 100k SH3": config_scene("C2"),
        "100k SH0": make_scene(100_000, 640, 480, sh_degree=0, seed=7),
        "30k SH0": make_scene(30_000, 640, 480, sh_degree=0, seed=5, px_min=1.0, px_max=8.0),
        "C3 500k SH0": config_scene("C3"),
    }
    native = ["g4r_forward_project", "g4r_forward_render", "g4r_wait_num_rendered", "g4r_backward", "g4r_geom_bytes",
              "g4r_image_bytes", "g4r_binning_bytes", "g4r_backward_scratch_bytes"]
    orig = {n: getattr(dgr._lib, n) for n in native}
    py_orig = {"_forward_impl": dgr._forward_impl, "_backward_impl": dgr._backward_impl}      # this package's Python between autograd and the C ABI
    out = {}
    for name, sc_cpu in cases.items():
        sc = sc_cpu.to(dev)
        row = {}
        for fresh in (True, False):
            loop(sc, dgr, 5, fresh)
            acc = {}
            for n in native:
                setattr(dgr._lib, n, Timed(orig[n], n, acc))
            for n, f in py_orig.items():
                setattr(dgr, n, (lambda f_, n_: (lambda *a, **k: Timed(lambda: f_(*a, **k), n_, acc)()))(f, n))
            wall, h = loop(sc, dgr, 50, fresh)
            for n in native:
                setattr(dgr._lib, n, orig[n])
            for n, f in py_orig.items():
                setattr(dgr, n, f)
            tag = "fresh_leaves" if fresh else "reused_leaves"
            row[tag] = {"wall_ms": round(wall, 4), "host_setup_ms": round(h[0], 4), "host_forward_ms": round(h[1], 4),
                        "host_loss_ms": round(h[2], 4), "host_backward_ms": round(h[3], 4),
                        "native_ms_per_iter": {k: round(v[0] / 50 * 1e3, 4) for k, v in acc.items()},
                        "native_calls_per_iter": {k: v[1] / 50 for k, v in acc.items()}}
        if ref is not None:
            loop(sc, ref, 5)
            wall, h = loop(sc, ref, 50)
            row["reference"] = {"wall_ms": round(wall, 4), "host_setup_ms": round(h[0], 4), "host_forward_ms": round(h[1], 4),
                                "host_loss_ms": round(h[2], 4), "host_backward_ms": round(h[3], 4)}
        out[name] = row
        print(name, json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "host_timeline.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
