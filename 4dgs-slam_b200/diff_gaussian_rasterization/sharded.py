"""Gaussian-sharded multi-GPU render (BASELINE.json config C4; no counterpart in the reference).

One process per GPU (``torch.distributed``, NCCL over NVLink).  Rank r owns a contiguous shard of the Gaussians (its
parameters, gradients and optimiser state never leave the rank) and the tiles ``t % world == r``.

forward   1. project the local shard                                   (g4r_project_only)
          2. all-gather the 48-byte splat records + radii              (NCCL all_gather, 52 B per Gaussian)
          3. count / scan / scatter / sort / composite the OWNED tiles over all records
                                                                         (g4r_count_tiles, g4r_forward_render)
          4. all-reduce(sum) the image planes (non-owned pixels are 0) (NCCL all_reduce, 20 B per pixel)
             reduce-scatter n_touched to the owning ranks
backward  5. composite backward of the owned tiles -> partial accumulators for ALL Gaussians (g4r_backward_composite)
          6. reduce-scatter(sum) the accumulators to the owning ranks  (NCCL reduce_scatter, 48 B per Gaussian)
          7. per-Gaussian backward of the local shard                  (g4r_backward_gaussians); all-reduce of dL/dtau

Global Gaussian ids are ``rank * Pmax + local index`` (shards padded to the largest one), which preserves the order of
the concatenated cloud, so every tile's sorted list -- and therefore every pixel -- is identical to the single-GPU
result.  The kernel calls go through a *backend* object so that the host logic (sharding, collectives, padding) can be
exercised with the gloo backend on CPU by the tests, which inject a CPU backend; the product default is the CUDA library
and there is no fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import (_BackwardIO, _ForwardOut, _check, _context, _dev_f32, _lib, _make_frame, _make_gaussians, _ptr,
               GaussianRasterizationSettings)

__all__ = ["ShardedGaussianRasterizer", "shard_bounds", "owned_tiles"]


def shard_bounds(P: int, world: int, rank: int):
    """Contiguous, balanced shard [lo, hi) of P Gaussians for `rank`."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owned_tiles(W: int, H: int, world: int, rank: int) -> torch.Tensor:
    """Boolean mask over tiles (row-major) owned by `rank`: interleaved ownership balances the per-tile load."""
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    return (torch.arange(tiles) % world) == rank


class CudaBackend:
    """The five kernel groups of the sharded render on the current CUDA device (C ABI of include/g4r.h)."""

    def project(self, rs, M, tile_rank, tile_world, means3D, opacities, sh, colors, scales, rots, cov, rec, radii, n_touched):
        dev = means3D.device
        keep = []
        with torch.cuda.device(dev):
            frame = _make_frame(rs, dev, M, keep)
            frame.tile_rank, frame.tile_world = tile_rank, tile_world
            g = _make_gaussians(int(means3D.shape[0]), means3D, opacities, sh, colors, scales, rots, cov)
            _check(_lib.g4r_project_only(ctypes.byref(frame), ctypes.byref(g), rec.data_ptr(), radii.data_ptr(), n_touched.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream))

    def render(self, rs, tile_rank, tile_world, P_all, rec_all, radii_all, n_touched_all, images, img_state, cap_hint):
        dev = rec_all.device
        keep = []
        with torch.cuda.device(dev):
            ctx = _context(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            frame = _make_frame(rs, dev, 0, keep)
            frame.tile_rank, frame.tile_world = tile_rank, tile_world
            g = _make_gaussians(P_all, rec_all, rec_all, None, None, None, None, None)
            _check(_lib.g4r_count_tiles(ctx, ctypes.byref(frame), P_all, radii_all.data_ptr(), rec_all.data_ptr(), img_state.data_ptr(), stream))
            out = _ForwardOut(images[0:3].data_ptr(), images[3:4].data_ptr(), images[4:5].data_ptr(), radii_all.data_ptr(),
                              n_touched_all.data_ptr())
            cap = cap_hint
            binning = torch.empty((_lib.g4r_binning_bytes(cap),), dtype=torch.uint8, device=dev)
            _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), rec_all.data_ptr(), img_state.data_ptr(),
                                           binning.data_ptr(), cap, ctypes.byref(out), stream))
            N = int(_lib.g4r_wait_num_rendered(ctx))
            if N < 0:
                _check(N)
            if N > cap:
                cap = N
                binning = torch.empty((_lib.g4r_binning_bytes(cap),), dtype=torch.uint8, device=dev)
                _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), rec_all.data_ptr(), img_state.data_ptr(),
                                               binning.data_ptr(), cap, ctypes.byref(out), stream))
        return binning, N

    def composite_backward(self, rs, tile_rank, tile_world, P_all, rec_all, img_state, binning, grad_color, grad_depth, acc_all):
        dev = rec_all.device
        keep = []
        with torch.cuda.device(dev):
            frame = _make_frame(rs, dev, 0, keep)
            frame.tile_rank, frame.tile_world = tile_rank, tile_world
            _check(_lib.g4r_backward_composite(ctypes.byref(frame), P_all, rec_all.data_ptr(), img_state.data_ptr(), binning.data_ptr(),
                                               grad_color.data_ptr(), grad_depth.data_ptr(), acc_all.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))

    def gaussian_backward(self, rs, M, means3D, sh, colors, scales, rots, cov, radii, rec, acc, grads: dict, tau):
        dev = means3D.device
        keep = []
        with torch.cuda.device(dev):
            frame = _make_frame(rs, dev, M, keep)
            g = _make_gaussians(int(means3D.shape[0]), means3D, means3D, sh, colors, scales, rots, cov)
            io = _BackwardIO(None, None, grads["means3D"].data_ptr(), grads["means2D"].data_ptr(), grads["opacities"].data_ptr(),
                             _ptr(grads.get("sh")), _ptr(grads.get("colors")), _ptr(grads.get("scales")), _ptr(grads.get("rots")),
                             _ptr(grads.get("cov")), tau.data_ptr())
            _check(_lib.g4r_backward_gaussians(ctypes.byref(frame), ctypes.byref(g), radii.data_ptr(), rec.data_ptr(), acc.data_ptr(),
                                               ctypes.byref(io), torch.cuda.current_stream(dev).cuda_stream))

    def image_state_bytes(self, W, H):
        return _lib.g4r_image_bytes(W, H)

    def geom_state_bytes(self, P):
        return _lib.g4r_geom_bytes(P)


_lib.g4r_project_only.restype = ctypes.c_int
_lib.g4r_project_only.argtypes = [ctypes.c_void_p] * 6
_lib.g4r_count_tiles.restype = ctypes.c_int
_lib.g4r_count_tiles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
_lib.g4r_backward_composite.restype = ctypes.c_int
_lib.g4r_backward_composite.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 7
_lib.g4r_backward_gaussians.restype = ctypes.c_int
_lib.g4r_backward_gaussians.argtypes = [ctypes.c_void_p] * 7

def _reduce_scatter_sum(out: torch.Tensor, inp: torch.Tensor, group) -> None:
    """reduce_scatter(sum); backends without it (gloo, used by the CPU tests) fall back to all_reduce + slice."""
    try:
        dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)
    except (RuntimeError, NotImplementedError):
        tmp = inp.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        n = out.shape[0]
        r = dist.get_rank(group)
        out.copy_(tmp[r * n:(r + 1) * n])


REC_FLOATS = 12       # 48-byte splat record
ACC_FLOATS = 12       # accumulator row (10 used)


class _ShardedRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, rs, group, backend):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = means3D.device
        H, W = int(rs.image_height), int(rs.image_width)
        P = int(means3D.shape[0])
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        means3D = _dev_f32(means3D, dev)
        opacities = _dev_f32(opacities, dev)
        sh = _dev_f32(sh, dev) if sh.numel() else sh
        colors_precomp = _dev_f32(colors_precomp, dev) if colors_precomp.numel() else colors_precomp
        scales = _dev_f32(scales, dev) if scales.numel() else scales
        rotations = _dev_f32(rotations, dev) if rotations.numel() else rotations
        cov3Ds_precomp = _dev_f32(cov3Ds_precomp, dev) if cov3Ds_precomp.numel() else cov3Ds_precomp
        M = int(sh.size(1)) if sh.numel() else 0

        # shard sizes -> padded global index space
        sizes = torch.tensor([P], dtype=torch.int64, device=dev)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes, group=group)
        Pmax = max(1, int(max(int(s.item()) for s in all_sizes)))
        P_all = world * Pmax

        # 1. local projection into a padded slab; padding rows keep radius 0 (invisible)
        rec_local = torch.zeros((Pmax, REC_FLOATS), **f32)
        radii_local = torch.zeros((Pmax,), **i32)
        ntouch_local = torch.zeros((Pmax,), **i32)
        # local geometry state: records first (the all-gathered slab), the per-Gaussian clamp bytes behind them
        geom_local = torch.zeros((max(Pmax * 48, backend.geom_state_bytes(max(P, 1))) + 256,), dtype=torch.uint8, device=dev)
        if P > 0:
            backend.project(rs, M, rank, world, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, geom_local,
                            radii_local, ntouch_local)
        rec_local = geom_local[: Pmax * 48].view(torch.float32).view(Pmax, REC_FLOATS)

        # 2. all-gather records and radii
        rec_all = torch.empty((P_all, REC_FLOATS), **f32)
        radii_all = torch.empty((P_all,), **i32)
        dist.all_gather_into_tensor(rec_all, rec_local.contiguous(), group=group)
        dist.all_gather_into_tensor(radii_all, radii_local, group=group)

        # 3. owned tiles: count / scan / scatter / sort / composite
        images = torch.zeros((5, H, W), **f32)                      # colour(3) depth(1) opacity(1); zeros outside owned tiles
        ntouch_all = torch.zeros((P_all,), **i32)
        img_state = torch.empty((backend.image_state_bytes(W, H),), dtype=torch.uint8, device=dev)
        cap = max(4096, int(1.5 * 4 * P_all / world))
        binning, N = backend.render(rs, rank, world, P_all, rec_all, radii_all, ntouch_all, images, img_state, cap)

        # 4. image all-reduce, n_touched reduce-scatter
        dist.all_reduce(images, op=dist.ReduceOp.SUM, group=group)
        ntouch_out = torch.empty((Pmax,), **i32)
        _reduce_scatter_sum(ntouch_out, ntouch_all, group)

        ctx.rs, ctx.group, ctx.backend, ctx.P, ctx.Pmax, ctx.M = rs, group, backend, P, Pmax, M
        ctx.opacities_shape = tuple(opacities.shape)
        ctx.binning = binning            # opaque to this layer (a byte tensor for the CUDA backend)
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local, rec_all, img_state)
        radii = radii_local[:P].clone()
        n_touched = ntouch_out[:P].clone()
        ctx.mark_non_differentiable(radii, n_touched)
        return images[0:3].clone(), radii, images[3:4].clone(), images[4:5].clone(), n_touched

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_opacity, grad_ntouched):
        means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local, rec_all, img_state = ctx.saved_tensors
        binning = ctx.binning
        rs, group, backend, P, Pmax, M = ctx.rs, ctx.group, ctx.backend, ctx.P, ctx.Pmax, ctx.M
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        P_all = world * Pmax
        grad_color = _dev_f32(grad_color, dev)
        grad_depth = _dev_f32(grad_depth, dev)

        # 5. + 6. partial accumulators of the owned tiles -> owners of the Gaussians
        acc_all = torch.empty((P_all, ACC_FLOATS), **f32)
        backend.composite_backward(rs, rank, world, P_all, rec_all, img_state, binning, grad_color, grad_depth, acc_all)
        acc_local = torch.empty((Pmax, ACC_FLOATS), **f32)
        _reduce_scatter_sum(acc_local, acc_all, group)

        # 7. per-Gaussian backward of the local shard
        grads = dict(means3D=torch.empty((P, 3), **f32), means2D=torch.empty((P, 3), **f32), opacities=torch.empty(ctx.opacities_shape, **f32))
        if sh.numel():
            grads["sh"] = torch.empty((P, M, 3), **f32)
        if colors_precomp.numel():
            grads["colors"] = torch.empty((P, 3), **f32)
        if scales.numel():
            grads["scales"] = torch.empty((P, 3), **f32)
            grads["rots"] = torch.empty((P, 4), **f32)
        if cov3Ds_precomp.numel():
            grads["cov"] = torch.empty((P, 6), **f32)
        tau = torch.zeros((8,), **f32)
        if P > 0:
            backend.gaussian_backward(rs, M, means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local,
                                      acc_local, grads, tau)
        dist.all_reduce(tau, op=dist.ReduceOp.SUM, group=group)
        needs = ctx.needs_input_grad
        return (grads["means3D"], grads["means2D"], grads.get("sh"), grads.get("colors"), grads["opacities"], grads.get("scales"),
                grads.get("rots"), grads.get("cov"), tau[3:6].view(1, -1) if needs[8] else None, tau[:3].view(1, -1) if needs[9] else None,
                None, None, None)


class ShardedGaussianRasterizer(torch.nn.Module):
    """Same call signature as ``GaussianRasterizer`` but every argument is the LOCAL shard of the Gaussians; returns the
    full image on every rank and the local ``radii`` / ``n_touched``.  The image gradients handed to backward must be
    identical on all ranks (every rank evaluates the loss on the full image)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings, group=None, backend=None):
        super().__init__()
        self.raster_settings = raster_settings
        self.group = group
        self.backend = backend if backend is not None else CudaBackend()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                theta=None, rho=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        e = lambda t: torch.Tensor([]) if t is None else t
        return _ShardedRasterize.apply(means3D, means2D, e(shs), e(colors_precomp), opacities, e(scales), e(rotations), e(cov3D_precomp),
                                       e(theta), e(rho), self.raster_settings, self.group, self.backend)
