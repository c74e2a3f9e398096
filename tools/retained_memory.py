"""Device memory held per live autograd context (BackEnd.map keeps ~10 rasterizer nodes alive until its single
backward(retain_graph=True), utils/slam_backend.py:657): peak allocated bytes while K forwards of workload C3 are retained, and
the bytes each context saves -- against what round 1 saved (the whole binning buffer: sorted ids + both pair arrays)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import runners  # noqa: E402
from tools.scenes import config_scene  # noqa: E402


def main():
    import diff_gaussian_rasterization as dgr
    dev = torch.device("cuda:0")
    sc = config_scene("C3").to(dev)
    rs = runners.settings_for(sc, dgr)
    K = 10
    leaf = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    for _ in range(2):          # warm the capacity hint
        runners.run_public_api(sc, dgr)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    loss = 0
    per_ctx = []
    for i in range(K):
        before = torch.cuda.memory_allocated()
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"])
        loss = loss + (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()
        torch.cuda.synchronize()
        per_ctx.append(torch.cuda.memory_allocated() - before)
    held = torch.cuda.memory_allocated() - base
    loss.backward(retain_graph=True)
    torch.cuda.synchronize()
    e = torch.Tensor([])
    st = dgr._forward_impl(sc.means3D, sc.shs, e, sc.opacities, sc.scales, sc.rotations, e, rs)[-1]
    cap = st["capacity"]
    rep = {"workload": "C3", "contexts": K, "held_MB_after_K_forwards": held / 1e6, "per_context_MB": sum(per_ctx) / K / 1e6,
           "peak_MB_incl_backward": (torch.cuda.max_memory_allocated() - base) / 1e6, "instance_capacity": cap,
           "saved_binning_MB_per_context": st["binning"].numel() / 1e6, "forward_only_sort_scratch_MB": st["sort_scratch"].numel() / 1e6,
           "round1_saved_binning_MB_per_context": (st["binning"].numel() + st["sort_scratch"].numel()) / 1e6,
           "saved_MB_per_context_vs_round1": st["sort_scratch"].numel() / 1e6}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "retained_memory.json"), "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
