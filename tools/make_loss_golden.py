"""Golden vectors for the fused SLAM losses, produced by the reference's OWN code: imports /root/reference/utils/slam_utils.py
(unmodified; matplotlib, which it imports for plotting only, is stubbed) in this container on the CPU, runs get_loss_tracking /
get_loss_mapping with torch autograd and stores inputs, loss and gradients in tests/golden/slam_loss.npz.

    python tools/make_loss_golden.py        # needs /root/reference; the committed fixture is what the tests read
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/utils/slam_utils.py"


class _Img:
    """viewpoint.original_image: the reference calls .cuda() on it (slam_utils.py:66,272); on this CPU box that is the identity."""
    def __init__(self, t):
        self.t = t

    def cuda(self):
        return self.t


def load_reference():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    spec = importlib.util.spec_from_file_location("ref_slam_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    cfg = {"Training": {"monocular": False, "rgb_boundary_threshold": 0.01, "alpha": 0.95}, "Results": {"save_dir": "/tmp/g4r_loss_golden"}}
    g = torch.Generator().manual_seed(123)
    h, w = 24, 32
    out = {}
    cases = [("track_uid0", "tracking", 0, True), ("track_uid3_motion", "tracking", 3, True), ("track_uid2_nomotion", "tracking", 2, False),
             ("map_motion", "mapping", 1, True), ("map_nomotion", "mapping", 1, False)]
    for name, mode, uid, with_motion in cases:
        image = torch.rand(3, h, w, generator=g).requires_grad_(True)
        depth = (0.2 + 3.0 * torch.rand(1, h, w, generator=g)).requires_grad_(True)
        opacity = torch.rand(1, h, w, generator=g) * 0.2 + 0.85
        gt_image = torch.rand(3, h, w, generator=g)
        gt_image[:, :3, :5] = 0.0                                 # below the rgb boundary threshold
        gt_depth = 0.2 + 3.0 * torch.rand(h, w, generator=g)
        gt_depth[5:8, 3:9] = 0.0                                  # invalid depth
        gt_depth[10, 10] = 5000.0                                 # beyond the tracking range, inside the mapping range
        vp = types.SimpleNamespace(uid=uid, original_image=_Img(gt_image), depth=gt_depth.numpy().copy(),
                                   exposure_a=torch.tensor([0.07], requires_grad=True), exposure_b=torch.tensor([-0.02], requires_grad=True),
                                   grad_mask=(torch.rand(1, h, w, generator=g) > 0.2),
                                   motion_mask=(torch.rand(h, w, generator=g) > 0.3) if with_motion else None)
        if mode == "tracking":
            loss = ref.get_loss_tracking(cfg, image, depth, opacity, vp, rm_dynamic=True, mask=None, save_img=False)
        else:
            loss = ref.get_loss_mapping(cfg, image, depth, vp, opacity, rm_dynamic=True)
        loss.backward()
        out.update({f"{name}/image": image.detach().numpy(), f"{name}/depth": depth.detach().numpy(), f"{name}/opacity": opacity.numpy(),
                    f"{name}/gt_image": gt_image.numpy(), f"{name}/gt_depth": gt_depth.numpy(), f"{name}/exposure": np.array([0.07, -0.02], np.float32),
                    f"{name}/grad_mask": vp.grad_mask.numpy(), f"{name}/uid": np.array(uid),
                    f"{name}/loss": loss.detach().numpy(), f"{name}/d_image": image.grad.numpy(), f"{name}/d_depth": depth.grad.numpy(),
                    f"{name}/d_exposure": np.array([vp.exposure_a.grad.item(), vp.exposure_b.grad.item()], np.float32)})
        if with_motion:
            out[f"{name}/motion_mask"] = vp.motion_mask.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "slam_loss.npz"), **out)
    print("wrote", len(cases), "cases:", [c[0] for c in cases])


if __name__ == "__main__":
    main()
