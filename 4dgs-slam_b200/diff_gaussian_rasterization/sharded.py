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
result.

``exchange="alltoall"`` replaces steps 2 and 6 by variable-size all-to-alls: rank r owns a contiguous strip of tile
rows, every splat record travels only to the ranks whose strip its tile rectangle touches (~1.3 ranks instead of all),
and the accumulator rows travel back the same way and are scatter-added at the owner.  Records arrive ordered by
(source rank, local index) = global id order, so ties in depth still resolve like on one GPU.  It moves ~6x fewer
bytes at 8 ranks, but on one NVSwitch node the extra host steps (split sizes, index build, index_add) cost more than
the bytes save (profiles/scaling/sharded_check_C4_x*.json: 1.99 ms vs 1.76 ms at 8 GPUs), so ``"allgather"`` -- the
variant described above, interleaved tile ownership -- is the default; ``"alltoall"`` is the one to pick when the
exchange crosses nodes.

The kernel calls go through a *backend* object so that the host logic (sharding, collectives, padding) can be
exercised with the gloo backend on CPU by the tests, which inject a CPU backend; the product default is the CUDA library
and there is no fallback.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import (_BackwardIO, _ForwardOut, _check, _context, _dev_f32, _lib, _make_frame, _make_gaussians, _ptr,
               GaussianRasterizationSettings)

__all__ = ["ShardedGaussianRasterizer", "shard_bounds", "owned_tiles", "strip_bounds"]


def shard_bounds(P: int, world: int, rank: int):
    """Contiguous, balanced shard [lo, hi) of P Gaussians for `rank`."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owned_tiles(W: int, H: int, world: int, rank: int) -> torch.Tensor:
    """Boolean mask over tiles (row-major) owned by `rank`: interleaved ownership balances the per-tile load."""
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    return (torch.arange(tiles) % world) == rank


def _set_owner(frame, owner) -> None:
    """owner = ("mod", rank, world) | ("rows", row_begin, row_end)"""
    frame.tile_rank, frame.tile_world, frame.tile_row_begin, frame.tile_row_end = 0, 1, 0, 0
    if owner[0] == "mod":
        frame.tile_rank, frame.tile_world = int(owner[1]), int(owner[2])
    else:
        frame.tile_row_begin, frame.tile_row_end = int(owner[1]), int(owner[2])


class CudaBackend:
    """The kernel groups of the sharded render on the current CUDA device (C ABI of include/g4r.h)."""

    def tile_rows(self, rs, P, radii, geom):
        dev = radii.device
        rows = torch.empty((max(P, 1), 2), dtype=torch.int32, device=dev)
        keep = []
        with torch.cuda.device(dev):
            frame = _make_frame(rs, dev, 0, keep)
            _check(_lib.g4r_tile_rows(ctypes.byref(frame), P, radii.data_ptr(), geom.data_ptr(), rows.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream))
        return rows[:P]

    def project(self, rs, M, means3D, opacities, sh, colors, scales, rots, cov, rec, radii, n_touched):
        dev = means3D.device
        keep = []
        with torch.cuda.device(dev):
            frame = _make_frame(rs, dev, M, keep)
            g = _make_gaussians(int(means3D.shape[0]), means3D, opacities, sh, colors, scales, rots, cov)
            _check(_lib.g4r_project_only(ctypes.byref(frame), ctypes.byref(g), rec.data_ptr(), radii.data_ptr(), n_touched.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream))

    def render(self, rs, owner, P_all, rec_all, radii_all, n_touched_all, images, img_state, cap_hint):
        dev = rec_all.device
        keep = []
        with torch.cuda.device(dev):
            ctx = _context(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            frame = _make_frame(rs, dev, 0, keep)
            _set_owner(frame, owner)
            g = _make_gaussians(P_all, rec_all, rec_all, None, None, None, None, None)
            _check(_lib.g4r_count_tiles(ctx, ctypes.byref(frame), P_all, radii_all.data_ptr(), rec_all.data_ptr(), img_state.data_ptr(), stream))
            out = _ForwardOut(images[0:3].data_ptr(), images[3:4].data_ptr(), images[4:5].data_ptr(), radii_all.data_ptr(),
                              n_touched_all.data_ptr())
            cap = cap_hint
            binning = torch.empty((_lib.g4r_binning_bytes(cap),), dtype=torch.uint8, device=dev)
            sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap),), dtype=torch.uint8, device=dev)
            _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), rec_all.data_ptr(), img_state.data_ptr(),
                                           binning.data_ptr(), sort_scratch.data_ptr(), cap, ctypes.byref(out), stream))
            N = int(_lib.g4r_wait_num_rendered(ctx))
            if N < 0:
                _check(N)
            if N > cap:
                cap = N
                binning = torch.empty((_lib.g4r_binning_bytes(cap),), dtype=torch.uint8, device=dev)
                sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap),), dtype=torch.uint8, device=dev)
                _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), rec_all.data_ptr(), img_state.data_ptr(),
                                               binning.data_ptr(), sort_scratch.data_ptr(), cap, ctypes.byref(out), stream))
        return binning, N

    def composite_backward(self, rs, owner, P_all, rec_all, img_state, binning, grad_color, grad_depth, acc_all):
        dev = rec_all.device
        keep = []
        with torch.cuda.device(dev):
            frame = _make_frame(rs, dev, 0, keep)
            _set_owner(frame, owner)
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


_lib.g4r_tile_rows.restype = ctypes.c_int
_lib.g4r_tile_rows.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 4
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
            backend.project(rs, M, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, geom_local, radii_local,
                            ntouch_local)
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
        binning, N = backend.render(rs, ("mod", rank, world), P_all, rec_all, radii_all, ntouch_all, images, img_state, cap)

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
        backend.composite_backward(rs, ("mod", rank, world), P_all, rec_all, img_state, binning, grad_color, grad_depth, acc_all)
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


def strip_bounds(tiles_y: int, world: int, rank: int):
    """Contiguous strip of tile rows [begin, end) owned by `rank`."""
    return (rank * tiles_y) // world, ((rank + 1) * tiles_y) // world


def _all_to_all_rows(send: torch.Tensor, send_counts, recv_counts, group) -> torch.Tensor:
    """Variable-size all-to-all of the rows of a 2-D tensor."""
    recv = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in send_counts], group=group)
    return recv


class _ShardedRasterizeA2A(torch.autograd.Function):
    """Strip ownership + variable-size all-to-all of splat records (forward) and accumulator rows (backward)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, rs, group, backend):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = means3D.device
        H, W = int(rs.image_height), int(rs.image_width)
        tiles_y = (H + 15) // 16
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

        # 1. local projection
        Pp = max(P, 1)
        geom_local = torch.zeros((max(Pp * 48, backend.geom_state_bytes(Pp)) + 256,), dtype=torch.uint8, device=dev)
        radii_local = torch.zeros((Pp,), **i32)
        ntouch_local = torch.zeros((Pp,), **i32)
        if P > 0:
            backend.project(rs, M, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, geom_local, radii_local,
                            ntouch_local)
        rec_local = geom_local[: Pp * 48].view(torch.float32).view(Pp, REC_FLOATS)

        # 2. destinations: every rank whose strip of tile rows intersects the splat's rectangle
        rows = backend.tile_rows(rs, P, radii_local, geom_local) if P > 0 else torch.zeros((0, 2), **i32)
        begins = torch.tensor([strip_bounds(tiles_y, world, r)[0] for r in range(world)], **i32).view(world, 1)
        ends = torch.tensor([strip_bounds(tiles_y, world, r)[1] for r in range(world)], **i32).view(world, 1)
        first, last = rows[:, 0].view(1, -1), rows[:, 1].view(1, -1)
        touch = (first < ends) & (last >= begins) & (last >= first)                  # (world, P)
        dest_src = torch.nonzero(touch)                                              # sorted by destination, then local index
        send_idx = dest_src[:, 1].contiguous()
        send_counts = touch.sum(dim=1).to(torch.int64)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        send_counts_l, recv_counts_l = send_counts.tolist(), recv_counts.tolist()

        # 3. exchange the 48-byte records; the radius rides in the record's spare slot
        send_rec = rec_local[send_idx].clone()
        send_rec[:, 11] = radii_local[send_idx].view(torch.float32) if send_idx.numel() else send_rec[:, 11]
        rec_recv = _all_to_all_rows(send_rec, send_counts_l, recv_counts_l, group)
        P_recv = int(rec_recv.shape[0])
        rec_work = rec_recv if P_recv > 0 else torch.zeros((1, REC_FLOATS), **f32)
        radii_recv = rec_work[:, 11].contiguous().view(torch.int32)

        # 4. owned strip: count / scan / scatter / sort / composite over the received records
        rb, re_ = strip_bounds(tiles_y, world, rank)
        images = torch.zeros((5, H, W), **f32)
        ntouch_recv = torch.zeros((max(P_recv, 1),), **i32)
        img_state = torch.empty((backend.image_state_bytes(W, H),), dtype=torch.uint8, device=dev)
        cap = max(4096, int(6 * max(P_recv, 1)))
        binning, N = backend.render(rs, ("rows", rb, re_), P_recv, rec_work, radii_recv, ntouch_recv, images, img_state, cap)

        # 5. image all-reduce (every pixel has exactly one owner), n_touched back to the owners of the Gaussians
        dist.all_reduce(images, op=dist.ReduceOp.SUM, group=group)
        ntouch_back = _all_to_all_rows(ntouch_recv[:P_recv].view(-1, 1), recv_counts_l, send_counts_l, group).view(-1)
        if send_idx.numel():
            ntouch_local.index_add_(0, send_idx, ntouch_back)

        ctx.rs, ctx.group, ctx.backend, ctx.P, ctx.M, ctx.P_recv = rs, group, backend, P, M, P_recv
        ctx.counts = (send_counts_l, recv_counts_l)
        ctx.owner = ("rows", rb, re_)
        ctx.opacities_shape = tuple(opacities.shape)
        ctx.binning = binning
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local, rec_work, img_state, send_idx)
        radii = radii_local[:P].clone()
        n_touched = ntouch_local[:P].clone()
        ctx.mark_non_differentiable(radii, n_touched)
        return images[0:3].clone(), radii, images[3:4].clone(), images[4:5].clone(), n_touched

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_opacity, grad_ntouched):
        means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local, rec_work, img_state, send_idx = ctx.saved_tensors
        rs, group, backend, P, M, P_recv = ctx.rs, ctx.group, ctx.backend, ctx.P, ctx.M, ctx.P_recv
        send_counts_l, recv_counts_l = ctx.counts
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        grad_color = _dev_f32(grad_color, dev)
        grad_depth = _dev_f32(grad_depth, dev)

        # composite backward of the owned strip -> one accumulator row per received record -> back to the owners
        acc_recv = torch.zeros((max(P_recv, 1), ACC_FLOATS), **f32)
        if P_recv > 0:
            backend.composite_backward(rs, ctx.owner, P_recv, rec_work, img_state, ctx.binning, grad_color, grad_depth, acc_recv)
        acc_back = _all_to_all_rows(acc_recv[:P_recv], recv_counts_l, send_counts_l, group)
        acc_local = torch.zeros((max(P, 1), ACC_FLOATS), **f32)
        if send_idx.numel():
            acc_local.index_add_(0, send_idx, acc_back)

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

    def __init__(self, raster_settings: GaussianRasterizationSettings, group=None, backend=None, exchange: str = "allgather"):
        super().__init__()
        if exchange not in ("alltoall", "allgather"):
            raise ValueError("exchange must be 'alltoall' or 'allgather'")
        self.raster_settings = raster_settings
        self.group = group
        self.backend = backend if backend is not None else CudaBackend()
        self.exchange = exchange

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                theta=None, rho=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        e = lambda t: torch.Tensor([]) if t is None else t
        fn = _ShardedRasterizeA2A if self.exchange == "alltoall" else _ShardedRasterize
        return fn.apply(means3D, means2D, e(shs), e(colors_precomp), opacities, e(scales), e(rotations), e(cov3D_precomp),
                        e(theta), e(rho), self.raster_settings, self.group, self.backend)
