"""Fused RGB-D losses of the two SLAM loops (opt-in extension).

``slam_loss(mode, image, depth, opacity, gt_image, gt_depth, exposure_a, exposure_b, ...)`` returns the scalar that the
reference's ``get_loss_tracking`` (mode "tracking"; utils/slam_utils.py:57-173, RGB-D branch) or ``get_loss_mapping``
(mode "mapping"; utils/slam_utils.py:252-364, RGB-D static non-split branch) returns, computed together with its gradients
w.r.t. the rendered colour / depth and the two exposure parameters by ONE kernel (csrc/loss.cu, C ABI ``g4r_slam_loss``)
instead of ~15 image-sized torch kernels forward and as many backward.  Gradients flow to ``image``, ``depth``,
``exposure_a`` and ``exposure_b``; ``opacity`` is a constant weight exactly like in the reference (the rasterizer drops the
gradient of its opacity output, DGR/diff_gaussian_rasterization/__init__.py:108).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _check, _dev_f32, _lib, _ptr

__all__ = ["slam_loss"]


class _LossIn(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("mode", ctypes.c_int32), ("alpha", ctypes.c_float),
                ("rgb_boundary_threshold", ctypes.c_float), ("image", ctypes.c_void_p), ("depth", ctypes.c_void_p),
                ("opacity", ctypes.c_void_p), ("gt_image", ctypes.c_void_p), ("gt_depth", ctypes.c_void_p), ("exposure_a", ctypes.c_void_p),
                ("exposure_b", ctypes.c_void_p), ("motion_mask", ctypes.c_void_p), ("grad_mask", ctypes.c_void_p)]


_lib.g4r_slam_loss.restype = ctypes.c_int
_lib.g4r_slam_loss.argtypes = [ctypes.POINTER(_LossIn)] + [ctypes.c_void_p] * 5


def _mask_u8(m: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if m is None:
        return None
    if m.dtype is torch.bool:
        m = m.view(torch.uint8) if m.is_contiguous() else m.contiguous().view(torch.uint8)
    elif m.dtype is not torch.uint8:
        m = (m != 0).view(torch.uint8)
    return m.to(device).contiguous()


class _SlamLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, exposure_a, exposure_b, opacity, gt_image, gt_depth, motion_mask, grad_mask, mode, alpha, thr):
        if not image.is_cuda:
            raise RuntimeError("slam_loss: tensors must live on a CUDA device; there is no CPU path")
        dev = image.device
        H, W = int(image.shape[-2]), int(image.shape[-1])
        image_c, depth_c = _dev_f32(image, dev), _dev_f32(depth, dev)
        gt_image_c, gt_depth_c = _dev_f32(gt_image, dev), _dev_f32(gt_depth, dev)
        opacity_c = _dev_f32(opacity, dev) if opacity is not None else None
        if mode == 0 and opacity_c is None:
            raise RuntimeError("slam_loss: the tracking loss needs the rendered opacity")
        ea = _dev_f32(exposure_a, dev) if exposure_a is not None else None
        eb = _dev_f32(exposure_b, dev) if exposure_b is not None else None
        mm, gm = _mask_u8(motion_mask, dev), _mask_u8(grad_mask, dev)
        d_image = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        d_depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        out = torch.empty((12,), dtype=torch.float32, device=dev)               # [0:4] results, [4:12] scratch
        arg = _LossIn(W, H, int(mode), float(alpha), float(thr), image_c.data_ptr(), depth_c.data_ptr(), _ptr(opacity_c), gt_image_c.data_ptr(),
                      gt_depth_c.data_ptr(), _ptr(ea), _ptr(eb), _ptr(mm), _ptr(gm))
        with torch.cuda.device(dev):
            _check(_lib.g4r_slam_loss(ctypes.byref(arg), d_image.data_ptr(), d_depth.data_ptr(), out.data_ptr(), out.data_ptr() + 16,
                                      torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(d_image, d_depth, out)
        ctx.shapes = (None if exposure_a is None else tuple(exposure_a.shape), None if exposure_b is None else tuple(exposure_b.shape))
        return out[0]

    @staticmethod
    def backward(ctx, g):
        d_image, d_depth, out = ctx.saved_tensors
        needs = ctx.needs_input_grad
        sa, sb = ctx.shapes
        return (d_image * g if needs[0] else None, d_depth * g if needs[1] else None,
                (out[1] * g).reshape(sa) if needs[2] and sa is not None else None,
                (out[2] * g).reshape(sb) if needs[3] and sb is not None else None,
                None, None, None, None, None, None, None, None)


def slam_loss(mode: str, image, depth, gt_image, gt_depth, opacity=None, exposure_a=None, exposure_b=None, motion_mask=None,
              grad_mask=None, alpha: float = 0.95, rgb_boundary_threshold: float = 0.01) -> torch.Tensor:
    """`mode` = "tracking" | "mapping".  `motion_mask` / `grad_mask`: bool or uint8 [H,W] (or broadcastable [1,H,W]) or None; pass
    `motion_mask=None` where the reference skips it (tracking: `viewpoint.uid == 0`)."""
    m = {"tracking": 0, "mapping": 1}[mode]
    return _SlamLoss.apply(image, depth, exposure_a, exposure_b, opacity, gt_image, gt_depth, motion_mask, grad_mask, m, alpha,
                           rgb_boundary_threshold)
