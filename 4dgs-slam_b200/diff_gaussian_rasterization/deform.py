"""Control-node warp of the deformation step in one kernel each way (opt-in extension).

``control_node_warp(x, nodes, log_radius, weight_logit, node_attrs, motion_mask, K=3, ...)`` returns what
``ControlNodeWarp.forward`` (utils/time_utils.py:1192-1275) returns after the node MLP has produced ``node_attrs`` -- the dict
``{'d_xyz', 'd_rotation', 'd_scaling'}`` that ``render(..., dx=, dr=, ds=)`` consumes (utils/slam_backend.py:361-372) -- including
``cal_nn_weight`` (:981-1015), so ``pytorch3d.ops.knn_points`` is not needed.  csrc/warp.cu, C ABI ``g4r_warp_forward`` /
``g4r_warp_backward``.  Gradients flow to ``log_radius`` (``_node_radius``), ``weight_logit`` (``_node_weight``) and the node
attributes ``d_xyz, d_rotation, d_scaling, local_rotation``; ``x``, the node positions and ``motion_mask`` are constants exactly
like in the reference (``x.detach()`` :1196, ``self.nodes[..., :3].detach()`` :994)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _check, _dev_f32, _lib, _ptr

__all__ = ["control_node_warp"]


class _WarpIn(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "M", "K", "node_stride", "d_rot_as_res", "local_frame")] + \
               [(n, ctypes.c_void_p) for n in ("x", "nodes", "log_radius", "weight_logit", "d_xyz", "d_rotation", "d_scaling", "local_rotation",
                                               "motion_mask")]


_lib.g4r_warp_forward.restype = ctypes.c_int
_lib.g4r_warp_forward.argtypes = [ctypes.POINTER(_WarpIn)] + [ctypes.c_void_p] * 7
_lib.g4r_warp_scratch_bytes.restype = ctypes.c_size_t
_lib.g4r_warp_scratch_bytes.argtypes = [ctypes.c_int32]
_lib.g4r_warp_backward.restype = ctypes.c_int
_lib.g4r_warp_backward.argtypes = [ctypes.POINTER(_WarpIn)] + [ctypes.c_void_p] * 14


class _ControlNodeWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_radius, weight_logit, d_xyz, d_rotation, d_scaling, local_rotation, x, nodes, motion_mask, K, d_rot_as_res, local_frame):
        if not x.is_cuda:
            raise RuntimeError("control_node_warp: tensors must live on a CUDA device; there is no CPU path")
        dev = x.device
        x_c, nodes_c = _dev_f32(x.detach(), dev), _dev_f32(nodes.detach(), dev)
        N, M = int(x_c.shape[0]), int(nodes_c.shape[0])
        lr = _dev_f32(log_radius.detach(), dev).reshape(-1)
        wl = _dev_f32(weight_logit.detach(), dev).reshape(-1) if weight_logit is not None else None
        dx, dr, ds = _dev_f32(d_xyz.detach(), dev), _dev_f32(d_rotation.detach(), dev), _dev_f32(d_scaling.detach(), dev)
        lrot = _dev_f32(local_rotation.detach(), dev) if (local_frame and local_rotation is not None) else None
        if local_frame and lrot is None:
            raise RuntimeError("control_node_warp: local_frame=True needs node_attrs['local_rotation']")
        mm = _dev_f32(motion_mask.detach(), dev).reshape(-1) if motion_mask is not None else None
        if mm is not None and mm.numel() != N:
            raise RuntimeError(f"control_node_warp: motion_mask has {mm.numel()} entries for {N} points")
        for name, t, cols in (("d_xyz", dx, 3), ("d_rotation", dr, 4), ("d_scaling", ds, 3)):
            if tuple(t.shape) != (M, cols):
                raise RuntimeError(f"control_node_warp: {name} must have shape [{M}, {cols}], got {tuple(t.shape)}")
        if lr.numel() != M or (wl is not None and wl.numel() != M):
            raise RuntimeError(f"control_node_warp: log_radius / weight_logit must have {M} entries")
        arg = _WarpIn(N, M, int(K), int(nodes_c.shape[1]), int(bool(d_rot_as_res)), int(bool(local_frame)), x_c.data_ptr(), nodes_c.data_ptr(),
                      lr.data_ptr(), _ptr(wl), dx.data_ptr(), dr.data_ptr(), ds.data_ptr(), _ptr(lrot), _ptr(mm))
        translate = torch.empty((N, 3), dtype=torch.float32, device=dev)
        rotation = torch.empty((N, 4), dtype=torch.float32, device=dev)
        scale = torch.empty((N, 3), dtype=torch.float32, device=dev)
        nn_idx = torch.empty((N, K), dtype=torch.int32, device=dev)
        nn_dist = torch.empty((N, K), dtype=torch.float32, device=dev)
        nn_weight = torch.empty((N, K), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _check(_lib.g4r_warp_forward(ctypes.byref(arg), translate.data_ptr(), rotation.data_ptr(), scale.data_ptr(), nn_idx.data_ptr(),
                                         nn_dist.data_ptr(), nn_weight.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(x_c, nodes_c, lr, wl, dx, dr, ds, lrot, mm, nn_idx, nn_dist, nn_weight)
        ctx.cfg = (int(K), int(bool(d_rot_as_res)), int(bool(local_frame)))
        ctx.shapes = (tuple(log_radius.shape), None if weight_logit is None else tuple(weight_logit.shape))
        ctx.mark_non_differentiable(nn_idx, nn_dist, nn_weight)
        return translate, rotation, scale, nn_weight, nn_dist, nn_idx

    @staticmethod
    def backward(ctx, g_translate, g_rotation, g_scale, *_):
        x_c, nodes_c, lr, wl, dx, dr, ds, lrot, mm, nn_idx, nn_dist, nn_weight = ctx.saved_tensors
        K, res, local = ctx.cfg
        dev = x_c.device
        N, M = int(x_c.shape[0]), int(nodes_c.shape[0])

        def grad_in(t, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if t is None else _dev_f32(t, dev)

        def grad_out(want, shape):
            return torch.empty(shape, dtype=torch.float32, device=dev) if want else None

        gT, gR, gS = grad_in(g_translate, (N, 3)), grad_in(g_rotation, (N, 4)), grad_in(g_scale, (N, 3))
        needs = ctx.needs_input_grad
        g_lr, g_wl = grad_out(needs[0], (M,)), grad_out(needs[1] and wl is not None, (M,))
        g_dx, g_dr, g_ds = grad_out(needs[2], (M, 3)), grad_out(needs[3], (M, 4)), grad_out(needs[4], (M, 3))
        g_lrot = grad_out(needs[5] and lrot is not None, (M, 4))
        arg = _WarpIn(N, M, K, int(nodes_c.shape[1]), res, local, x_c.data_ptr(), nodes_c.data_ptr(), lr.data_ptr(), _ptr(wl), dx.data_ptr(),
                      dr.data_ptr(), ds.data_ptr(), _ptr(lrot), _ptr(mm))
        scratch = torch.empty(int(_lib.g4r_warp_scratch_bytes(M)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _check(_lib.g4r_warp_backward(ctypes.byref(arg), nn_idx.data_ptr(), nn_dist.data_ptr(), nn_weight.data_ptr(), gT.data_ptr(), gR.data_ptr(),
                                          gS.data_ptr(), _ptr(g_dx), _ptr(g_dr), _ptr(g_ds), _ptr(g_lrot), _ptr(g_lr), _ptr(g_wl),
                                          scratch.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        s_lr, s_wl = ctx.shapes
        return (g_lr.reshape(s_lr) if g_lr is not None else None, g_wl.reshape(s_wl) if g_wl is not None else None, g_dx, g_dr, g_ds, g_lrot,
                None, None, None, None, None, None)


def control_node_warp(x: torch.Tensor, nodes: torch.Tensor, log_radius: torch.Tensor, weight_logit: Optional[torch.Tensor], node_attrs: dict,
                      motion_mask: Optional[torch.Tensor] = None, K: int = 3, d_rot_as_res: bool = True, local_frame: bool = True,
                      return_nn: bool = False) -> dict:
    """``x`` [N,3] positions of the dynamic Gaussians, ``nodes`` [M, 3 + hyper_dim] (``ControlNodeWarp.nodes``), ``log_radius`` [M]
    (``_node_radius``), ``weight_logit`` [M,1] (``_node_weight``) or None, ``node_attrs`` = the node MLP's output dict
    (``node_deform``, :1038-1050: 'd_xyz' [M,3], 'd_rotation' [M,4], 'd_scaling' [M,3], 'local_rotation' [M,4]), ``motion_mask`` [N,1]
    or None.  Returns {'d_xyz', 'd_rotation', 'd_scaling', 'd_opacity': None, 'd_color': None} (+ 'nn_weight', 'nn_dist', 'nn_idx'
    when ``return_nn``: what cal_nn_weight returns)."""
    if not 1 <= int(K) <= 8:
        raise ValueError(f"control_node_warp: K = {K} is outside [1, 8]")
    t, r, s, nn_weight, nn_dist, nn_idx = _ControlNodeWarp.apply(log_radius, weight_logit, node_attrs["d_xyz"], node_attrs["d_rotation"],
                                                                 node_attrs["d_scaling"], node_attrs.get("local_rotation"), x, nodes, motion_mask,
                                                                 int(K), bool(d_rot_as_res), bool(local_frame))
    out = {"d_xyz": t, "d_rotation": r, "d_scaling": s, "d_opacity": None, "d_color": None}
    if return_nn:
        out.update(nn_weight=nn_weight, nn_dist=nn_dist, nn_idx=nn_idx.long())
    return out
