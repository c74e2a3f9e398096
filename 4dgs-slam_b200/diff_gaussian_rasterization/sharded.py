"""Gaussian-sharded multi-GPU render (BASELINE.json config C4; no counterpart in the reference).

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch).  Rank r owns a contiguous shard of the Gaussians
(its parameters, gradients and optimiser state never leave the rank) and the contiguous strip of tile rows
``[r * tiles_y / world, (r + 1) * tiles_y / world)``.

forward   1. project the local shard                                              g4r_project_only
          2. pack, per destination rank, the 48-byte splat records whose tile rectangle touches that rank's strip
             (stable compaction on the device, fixed-capacity slabs)              g4r_shard_pack
          3. all-to-all of the slabs (equal splits: no host-side sizes; each slab's header row carries its count)
          4. bin / sort / composite the OWNED strip over the received records, straight into an all-gather send buffer
                                                                                  g4r_shard_unpack, g4r_count_tiles, g4r_forward_render
          5. ONE all-gather whose payload is [image strip | n_touched of the received records | row of the count matrix]
             -> full image on every rank (g4r_shard_assemble), n_touched summed at the owners (g4r_shard_gather)
          (a 4-byte all-reduce(max) of the pair counts sits in front of step 3: the overflow check of the slab capacity)
backward  6. composite backward of the owned strip -> one accumulator row per received record   g4r_backward_composite
          7. reverse all-to-all of the rows; every owner sums the rows of its Gaussians         g4r_shard_gather
          8. per-Gaussian backward of the local shard; all-reduce of the 6 pose-gradient floats g4r_backward_gaussians

Nothing on this path reads a size back to the host before the frame is fully enqueued: the slab capacity per (source,
destination) pair is speculative (high-water mark of the count matrix x 1.25, identical on all ranks because every rank sees
the whole matrix) and the matrix is inspected after the last enqueue; a pair that outgrew the capacity makes ALL ranks redo
the frame with a larger one -- the same scheme the single-GPU path uses for its instance capacity.

Bit-exactness: the pack is a stable compaction, so inside a source's slab records keep their local order, and slabs are laid
out by source rank: the index of a received record is monotone in the global Gaussian id.  A tile's list is sorted by
(depth bits, index) -- the order of the single-GPU render -- so every pixel is bit-identical to it.

The kernel calls go through a *backend* object so that the host logic (strips, slabs, collectives, redo) can be exercised
with the gloo backend on CPU by the tests, which inject a CPU backend built on the oracle; the product default is the CUDA
library and there is no fallback.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch
import torch.distributed as dist

from . import (_BackwardIO, _ForwardOut, _check, _context, _dev_f32, _lib, _make_frame, _make_gaussians, _ptr,
               GaussianRasterizationSettings)

__all__ = ["ShardedGaussianRasterizer", "shard_bounds", "strip_bounds", "collective_bytes"]

REC_FLOATS = 12       # 48-byte splat record
ACC_FLOATS = 12       # accumulator row (10 used)
PLANES = 5            # colour (3), depth, opacity


def shard_bounds(P: int, world: int, rank: int):
    """Contiguous, balanced shard [lo, hi) of P Gaussians for `rank`."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def strip_bounds(tiles_y: int, world: int, rank: int):
    """Contiguous strip of tile rows [begin, end) owned by `rank` (same arithmetic as csrc/shard.cu make_geom)."""
    return (rank * tiles_y) // world, ((rank + 1) * tiles_y) // world


def _strip_rows(H: int, world: int):
    """(tiles_y, max strip height in pixels)."""
    ty = (H + 15) // 16
    return ty, 16 * max(strip_bounds(ty, world, r)[1] - strip_bounds(ty, world, r)[0] for r in range(world))


def collective_bytes(W: int, H: int, world: int, cap: int) -> dict:
    """Bytes every rank SENDS per frame in each collective (what bench.py reports next to the sharded timing)."""
    _, maxh = _strip_rows(H, world)
    rows = cap + 1
    return {"all_reduce_max_pair_count": 4, "all_to_all_records": world * rows * REC_FLOATS * 4,
            "all_gather_strip_ntouched_counts": (PLANES * maxh * W + world * rows + world) * 4,
            "all_to_all_accumulators": world * rows * ACC_FLOATS * 4, "all_reduce_pose_gradient": 32}


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
_lib.g4r_shard_scratch_bytes.restype = ctypes.c_size_t
_lib.g4r_shard_scratch_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
_lib.g4r_shard_pack.restype = ctypes.c_int
_lib.g4r_shard_pack.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64] + [ctypes.c_void_p] * 5
_lib.g4r_shard_unpack.restype = ctypes.c_int
_lib.g4r_shard_unpack.argtypes = [ctypes.c_int32, ctypes.c_int64] + [ctypes.c_void_p] * 4
_lib.g4r_shard_fetch_counts.restype = ctypes.c_int
_lib.g4r_shard_fetch_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p]
_lib.g4r_shard_wait_counts.restype = ctypes.c_int
_lib.g4r_shard_wait_counts.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
_lib.g4r_shard_gather.restype = ctypes.c_int
_lib.g4r_shard_gather.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
_lib.g4r_shard_assemble.restype = ctypes.c_int
_lib.g4r_shard_assemble.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64] + [ctypes.c_void_p] * 3


class CudaBackend:
    """The kernel groups of the sharded render on the current CUDA device (C ABI of include/g4r.h).  `fr` is the per-frame
    bundle made by `frames()` (the ctypes camera structs are built once per frame, not once per kernel group)."""

    def geom_state_bytes(self, P):
        return _lib.g4r_geom_bytes(P)

    def image_state_bytes(self, W, H):
        return _lib.g4r_image_bytes(W, H)

    def frames(self, rs, dev, M, rows):
        keep = []
        with torch.cuda.device(dev):
            full = _make_frame(rs, dev, M, keep)
            strip = _make_frame(rs, dev, 0, keep)
        strip.tile_rank, strip.tile_world, strip.tile_row_begin, strip.tile_row_end = 0, 1, int(rows[0]), int(rows[1])
        return dict(full=full, strip=strip, keep=keep, dev=dev, stream=torch.cuda.current_stream(dev).cuda_stream)

    def project(self, fr, means3D, opacities, sh, colors, scales, rots, cov, geom, radii, n_touched):
        with torch.cuda.device(fr["dev"]):
            g = _make_gaussians(int(means3D.shape[0]), means3D, opacities, sh, colors, scales, rots, cov)
            _check(_lib.g4r_project_only(ctypes.byref(fr["full"]), ctypes.byref(g), geom.data_ptr(), radii.data_ptr(), n_touched.data_ptr(),
                                         fr["stream"]))

    def pack(self, fr, P, radii, geom, world, cap, send_slab, payload, counts_offset, slots):
        """`counts` (this rank's row of the count matrix) is written into payload[counts_offset : counts_offset + world]."""
        scratch = torch.empty((_lib.g4r_shard_scratch_bytes(P, world),), dtype=torch.uint8, device=fr["dev"])
        with torch.cuda.device(fr["dev"]):
            _check(_lib.g4r_shard_pack(ctypes.byref(fr["full"]), P, radii.data_ptr(), geom.data_ptr(), world, cap, send_slab.data_ptr(),
                                       payload.data_ptr() + 4 * counts_offset, slots.data_ptr(), scratch.data_ptr(), fr["stream"]))

    def render_strip(self, fr, rows, world, cap, recv_slab, payload, strip_elems, maxh, W, img_state, cap_hint):
        """Unpack + bin + sort + composite of the strip `rows` = (begin, end) over the world*(cap+1) received slots.  `payload` is
        this rank's all-gather send buffer: the strip [5, maxh, W] at element 0, the n_touched slots behind it.  Returns (binning, N)."""
        dev = fr["dev"]
        P_all = world * (cap + 1)
        radii_all = torch.empty((P_all,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ctx = _context(dev)
            stream = fr["stream"]
            frame = fr["strip"]
            strip_ptr = payload.data_ptr()
            nt_ptr = strip_ptr + 4 * strip_elems
            _check(_lib.g4r_shard_unpack(world, cap, recv_slab.data_ptr(), radii_all.data_ptr(), nt_ptr, stream))
            g = _make_gaussians(P_all, recv_slab, recv_slab, None, None, None, None, None)
            _check(_lib.g4r_count_tiles(ctx, ctypes.byref(frame), P_all, radii_all.data_ptr(), recv_slab.data_ptr(), img_state.data_ptr(), stream))
            base = strip_ptr - 4 * (16 * int(rows[0])) * W              # pixel row y of the image = row y - 16*begin of the strip
            plane = maxh * W
            out = _ForwardOut(base, base + 4 * 3 * plane, base + 4 * 4 * plane, radii_all.data_ptr(), nt_ptr, plane)
            cap_n = cap_hint
            u8 = dict(dtype=torch.uint8, device=dev)
            binning = torch.empty((_lib.g4r_binning_bytes(cap_n),), **u8)
            sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap_n),), **u8)
            _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), recv_slab.data_ptr(), img_state.data_ptr(),
                                           binning.data_ptr(), sort_scratch.data_ptr(), cap_n, ctypes.byref(out), stream))
            N = int(_lib.g4r_wait_num_rendered(ctx))
            if N < 0:
                _check(N)
            if N > cap_n:
                cap_n = N
                binning = torch.empty((_lib.g4r_binning_bytes(cap_n),), **u8)
                sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap_n),), **u8)
                _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), recv_slab.data_ptr(), img_state.data_ptr(),
                                               binning.data_ptr(), sort_scratch.data_ptr(), cap_n, ctypes.byref(out), stream))
        return binning, N

    def assemble(self, fr, world, maxh, payload, rank_stride, images):
        with torch.cuda.device(fr["dev"]):
            _check(_lib.g4r_shard_assemble(ctypes.byref(fr["full"]), world, PLANES, maxh, rank_stride, payload.data_ptr(), images.data_ptr(),
                                           fr["stream"]))

    def fetch_scalar(self, fr, t: torch.Tensor):
        """Start the asynchronous copy of a 1-element int32 device tensor into the native context's pinned buffer; returns a
        function that waits for THAT copy only (an event recorded right behind it) and returns the value."""
        dev = fr["dev"]
        with torch.cuda.device(dev):
            ctx = _context(dev)
            _check(_lib.g4r_shard_fetch_counts(ctx, t.data_ptr(), 4, 0, 1, fr["stream"]))

        def wait():
            buf = (ctypes.c_int32 * 1)()
            _check(_lib.g4r_shard_wait_counts(ctx, 1, buf))
            return int(buf[0])
        return wait

    def gather(self, fr, P, world, cap, slots, acc_back=None, acc_stride=0, acc_local=None, nt_src=None, nt_offset=0, nt_stride=0, n_touched=None):
        """acc_back: [world, stride, 12] rows; nt_src: the all-gathered payload (flat), this rank's n_touched segment of rank d's
        payload starts at element nt_offset + d * nt_stride."""
        nt_ptr = None if nt_src is None else nt_src.data_ptr() + 4 * nt_offset
        with torch.cuda.device(fr["dev"]):
            _check(_lib.g4r_shard_gather(P, world, cap, slots.data_ptr(), _ptr(acc_back), acc_stride, _ptr(acc_local), nt_ptr, nt_stride,
                                         _ptr(n_touched), fr["stream"]))

    def composite_backward(self, fr, P_all, rec_all, img_state, binning, grad_color, grad_depth, acc_all):
        with torch.cuda.device(fr["dev"]):
            _check(_lib.g4r_backward_composite(ctypes.byref(fr["strip"]), P_all, rec_all.data_ptr(), img_state.data_ptr(), binning.data_ptr(),
                                               grad_color.data_ptr(), grad_depth.data_ptr(), acc_all.data_ptr(), fr["stream"]))

    def gaussian_backward(self, fr, means3D, sh, colors, scales, rots, cov, radii, geom, acc, grads: dict, tau):
        with torch.cuda.device(fr["dev"]):
            g = _make_gaussians(int(means3D.shape[0]), means3D, means3D, sh, colors, scales, rots, cov)
            io = _BackwardIO(None, None, grads["means3D"].data_ptr(), grads["means2D"].data_ptr(), grads["opacities"].data_ptr(),
                             _ptr(grads.get("sh")), _ptr(grads.get("colors")), _ptr(grads.get("scales")), _ptr(grads.get("rots")),
                             _ptr(grads.get("cov")), tau.data_ptr())
            _check(_lib.g4r_backward_gaussians(ctypes.byref(fr["full"]), ctypes.byref(g), radii.data_ptr(), geom.data_ptr(), acc.data_ptr(),
                                               ctypes.byref(io), fr["stream"]))


# per (device, W, H, world) state shared by the per-call rasterizer objects: slab capacity per rank pair and the instance capacity
_shard_lock = threading.Lock()
_shard_state: dict = {}


def _state(key) -> dict:
    with _shard_lock:
        st = _shard_state.get(key)
        if st is None:
            st = _shard_state[key] = dict(pair_hint=0, n_hint=0, P=None, Pmax=None, cap=None, redos=0)
        return st


def last_capacity(device, W: int, H: int, world: int):
    """Slab capacity (records per rank pair) the most recent frame of this shape ran with (bench.py reports bytes from it)."""
    return _state((str(torch.device(device)), W, H, world)).get("cap")


class _ShardedRasterize(torch.autograd.Function):
    """Strip ownership + fixed-capacity all-to-all of splat records (forward) and accumulator rows (backward).  Two
    collectives per direction: forward = all-to-all of the slabs + ONE all-gather whose payload is this rank's image strip,
    the n_touched of the records it received and its row of the count matrix; backward = reverse all-to-all of the
    accumulator rows + the 6-float pose-gradient all-reduce."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, rs, group, backend):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = means3D.device
        H, W = int(rs.image_height), int(rs.image_width)
        tiles_y, maxh = _strip_rows(H, world)
        if world > tiles_y:
            raise ValueError(f"{world} ranks but only {tiles_y} tile rows: a rank would own an empty strip")
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
        st = _state((str(dev), W, H, world))

        # shard sizes: the largest one bounds the slab capacity of the very first frame (cached until the model is resized)
        if st["P"] != P or st["Pmax"] is None:
            sizes = torch.tensor([P], dtype=torch.int64, device=dev)
            all_sizes = torch.empty((world,), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_sizes.view(-1), sizes.view(-1), group=group)
            st["Pmax"], st["P"] = max(1, int(all_sizes.max().item())), P
        Pmax = st["Pmax"]

        rb, re_ = strip_bounds(tiles_y, world, rank)
        fr = backend.frames(rs, dev, M, (rb, re_))
        # 1. local projection
        Pp = max(P, 1)
        geom_local = torch.empty((backend.geom_state_bytes(Pp),), dtype=torch.uint8, device=dev)
        radii_local = torch.zeros((Pp,), **i32) if P == 0 else torch.empty((Pp,), **i32)
        ntouch_local = torch.zeros((Pp,), **i32) if P == 0 else torch.empty((Pp,), **i32)
        if P > 0:
            backend.project(fr, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, geom_local, radii_local, ntouch_local)
        strip_elems = PLANES * maxh * W
        for _attempt in range(3):
            cap = min(Pmax, int(st["pair_hint"] * 1.25) + 256) if st["pair_hint"] > 0 else Pmax
            slab_rows = cap + 1                                             # + the header row that carries the count
            # all-gather payload of this rank: [image strip | n_touched of the received records | its row of the count matrix]
            nt_elems = world * slab_rows
            counts_offset = strip_elems + nt_elems
            payload_elems = (counts_offset + world + 3) // 4 * 4            # 16-byte multiples keep the strip assembly vectorised
            payload = torch.empty((payload_elems,), **f32)
            # 2. pack per destination (writes the slabs, their headers, this rank's counts and the slot table)
            send_slab = torch.empty((world, slab_rows, REC_FLOATS), **f32)
            slots = torch.empty((world, Pp), **i32)
            backend.pack(fr, P, radii_local, geom_local, world, cap, send_slab, payload, counts_offset, slots)
            # The largest count of any (source, destination) pair decides whether the slabs were big enough.  It is reduced
            # right here -- a 4-byte all-reduce in front of the all-to-all -- so that the host can look at it at the end of
            # forward without waiting for anything late in the stream (the composite is still running then).
            worst_dev = payload[counts_offset:counts_offset + world].view(torch.int32).max().view(1)
            dist.all_reduce(worst_dev, op=dist.ReduceOp.MAX, group=group)
            max_count = backend.fetch_scalar(fr, worst_dev)
            # 3. exchange the slabs
            recv_slab = torch.empty((world, slab_rows, REC_FLOATS), **f32)
            dist.all_to_all_single(recv_slab, send_slab, group=group)
            # 4. owned strip
            img_state = torch.empty((backend.image_state_bytes(W, H),), dtype=torch.uint8, device=dev)
            cap_n = int(st["n_hint"] * 1.25) + 4096 if st["n_hint"] > 0 else max(4096, 6 * cap)
            binning, N = backend.render_strip(fr, (rb, re_), world, cap, recv_slab, payload, strip_elems, maxh, W, img_state, cap_n)
            st["n_hint"] = max(N, int(st["n_hint"] * 0.95))
            # 5. one all-gather: strips -> full image on every rank, n_touched -> owners
            gathered = torch.empty((world * payload_elems,), **f32)
            dist.all_gather_into_tensor(gathered, payload, group=group)
            images = torch.empty((PLANES, H, W), **f32)
            backend.assemble(fr, world, maxh, gathered, payload_elems, images)
            if P > 0:
                backend.gather(fr, P, world, cap, slots, nt_src=gathered, nt_offset=strip_elems + rank * slab_rows, nt_stride=payload_elems,
                               n_touched=ntouch_local)
            # everything is enqueued: now look at the largest pair count (identical on all ranks -> identical decision)
            worst = int(max_count())
            st["pair_hint"] = max(worst, int(st["pair_hint"] * 0.95))
            st["cap"] = cap
            if worst <= cap:
                break
            st["redos"] += 1                 # some pair's slab was too small: every rank redoes the frame with the new hint
        else:
            raise RuntimeError("sharded render: slab capacity kept overflowing")

        ctx.rs, ctx.group, ctx.backend, ctx.P, ctx.M = rs, group, backend, P, M
        ctx.world, ctx.cap, ctx.fr = world, cap, fr
        ctx.opacities_shape = tuple(opacities.shape)
        ctx.binning = binning            # opaque to this layer (a byte tensor for the CUDA backend)
        ctx.gathered = gathered          # keeps the n_touched segment alive until the gather kernel has run (stream-ordered anyway)
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local, recv_slab, img_state, slots)
        radii = radii_local[:P]
        n_touched = ntouch_local[:P]
        ctx.mark_non_differentiable(radii, n_touched)
        return images[0:3], radii, images[3:4], images[4:5], n_touched

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_opacity, grad_ntouched):
        means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local, recv_slab, img_state, slots = ctx.saved_tensors
        group, backend, P, M, fr = ctx.group, ctx.backend, ctx.P, ctx.M, ctx.fr
        world, cap = ctx.world, ctx.cap
        ctx.gathered = None
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        grad_color = _dev_f32(grad_color, dev)
        grad_depth = _dev_f32(grad_depth, dev)
        slab_rows = cap + 1

        # 6. + 7. accumulator rows of the received records -> back to the owners -> summed per Gaussian
        acc_all = torch.empty((world, slab_rows, ACC_FLOATS), **f32)
        backend.composite_backward(fr, world * slab_rows, recv_slab, img_state, ctx.binning, grad_color, grad_depth, acc_all)
        acc_back = torch.empty((world, slab_rows, ACC_FLOATS), **f32)
        dist.all_to_all_single(acc_back, acc_all, group=group)
        acc_local = torch.empty((max(P, 1), ACC_FLOATS), **f32)
        if P > 0:
            backend.gather(fr, P, world, cap, slots, acc_back=acc_back, acc_stride=slab_rows, acc_local=acc_local)

        # 8. per-Gaussian backward of the local shard
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
            backend.gaussian_backward(fr, means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii_local, geom_local,
                                      acc_local, grads, tau)
        needs = ctx.needs_input_grad
        if needs[8] or needs[9]:
            dist.all_reduce(tau, op=dist.ReduceOp.SUM, group=group)
        return (grads["means3D"], grads["means2D"], grads.get("sh"), grads.get("colors"), grads["opacities"], grads.get("scales"),
                grads.get("rots"), grads.get("cov"), tau[3:6].view(1, -1) if needs[8] else None, tau[:3].view(1, -1) if needs[9] else None,
                None, None, None)


# ----------------------------------------------------------------------------------------------------------------------
# native runtime (csrc/shard_nccl.cu): the frame's kernels AND collectives are enqueued by three C calls per forward and one per
# backward.  Same algorithm, buffers and results as the Python-orchestrated path above; ~4x less host time per frame.
# ----------------------------------------------------------------------------------------------------------------------
class _ShardBuffers(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("cap", ctypes.c_int64), ("cap_n", ctypes.c_int64),
                ("geom_local", ctypes.c_void_p), ("radii_local", ctypes.c_void_p), ("n_touched_local", ctypes.c_void_p),
                ("send_slab", ctypes.c_void_p), ("recv_slab", ctypes.c_void_p), ("slots", ctypes.c_void_p), ("pack_scratch", ctypes.c_void_p),
                ("payload", ctypes.c_void_p), ("strip_elems", ctypes.c_int64), ("maxh", ctypes.c_int64), ("counts_offset", ctypes.c_int64),
                ("payload_elems", ctypes.c_int64), ("gathered", ctypes.c_void_p), ("images", ctypes.c_void_p), ("img_state", ctypes.c_void_p),
                ("binning", ctypes.c_void_p), ("sort_scratch", ctypes.c_void_p), ("radii_all", ctypes.c_void_p), ("worst", ctypes.c_void_p)]


_vp = ctypes.c_void_p
_lib.g4r_shard_buffers_size.restype = ctypes.c_int
_lib.g4r_shard_nccl_load.restype = ctypes.c_int
_lib.g4r_shard_nccl_load.argtypes = [ctypes.c_char_p]
_lib.g4r_shard_nccl_unique_id.restype = ctypes.c_int
_lib.g4r_shard_nccl_unique_id.argtypes = [_vp]
_lib.g4r_shard_comm_create.restype = ctypes.c_int
_lib.g4r_shard_comm_create.argtypes = [_vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_vp)]
_lib.g4r_shard_forward_a.restype = ctypes.c_int
_lib.g4r_shard_forward_a.argtypes = [_vp] * 7
_lib.g4r_shard_render_owned.restype = ctypes.c_int
_lib.g4r_shard_render_owned.argtypes = [_vp] * 4
_lib.g4r_shard_forward_wait.restype = ctypes.c_int
_lib.g4r_shard_forward_wait.argtypes = [_vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
_lib.g4r_shard_forward_b.restype = ctypes.c_int
_lib.g4r_shard_forward_b.argtypes = [_vp, _vp, ctypes.c_int32, _vp, _vp]
_lib.g4r_shard_backward.restype = ctypes.c_int
_lib.g4r_shard_backward.argtypes = [_vp] * 11 + [ctypes.c_int32, _vp]
if _lib.g4r_shard_buffers_size() != ctypes.sizeof(_ShardBuffers):
    raise ImportError("diff_gaussian_rasterization.sharded: G4RShardBuffers layout mismatch between libg4r.so and the Python bindings")

_comm_lock = threading.Lock()
_comms: dict = {}


def _nccl_path():
    """The NCCL library this process already has loaded (torch's)."""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl" in line:
                    return line.split()[-1]
    except OSError:
        pass
    return None


def native_comm(group, dev):
    """This library's own NCCL communicator over the ranks of `group` (created once per (group, device); collective)."""
    key = (id(group) if group is not None else 0, str(dev))
    with _comm_lock:
        h = _comms.get(key)
    if h is not None:
        return h
    path = _nccl_path()
    _check(_lib.g4r_shard_nccl_load(path.encode() if path else None))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_uint8 * 128)()
        _check(_lib.g4r_shard_nccl_unique_id(buf))
        ident = torch.tensor(list(buf), dtype=torch.uint8)
    ident = ident.to(dev)
    dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = (ctypes.c_uint8 * 128)(*ident.cpu().tolist())
    out = ctypes.c_void_p()
    with torch.cuda.device(dev):
        _check(_lib.g4r_shard_comm_create(raw, rank, world, ctypes.byref(out)))
    with _comm_lock:
        _comms[key] = out.value
    return out.value


def _settle(st: dict, nctx) -> None:
    """Deferred mode: read the two early events of the last frame rendered with this state (they fired long ago, so this does
    not block), update the capacity hints, and raise if that frame had outgrown its buffers -- its outputs were invalid."""
    pend = st.get("pending")
    if pend is None:
        return
    st["pending"] = None
    N, worst = ctypes.c_int64(), ctypes.c_int64()
    _check(_lib.g4r_shard_forward_wait(nctx, ctypes.byref(N), ctypes.byref(worst)))
    N, worst = int(N.value), int(worst.value)
    st["pair_hint"] = max(worst, int(st["pair_hint"] * 0.95))
    st["n_hint"] = max(N, int(st["n_hint"] * 0.95))
    if worst > pend[0] or N > pend[1]:
        st["redos"] += 1
        raise RuntimeError(f"sharded render (deferred check): the previous frame outgrew its buffers (pair count {worst} > {pend[0]} or "
                           f"instances {N} > {pend[1]}); its outputs were invalid.  The capacities have been raised: render that frame again")


def _sticky(st: dict, key: str, need: int, limit: int) -> int:
    """Capacity with hysteresis: grows to 1.25 x need (rounded up to 4096) when the current one is too small and only shrinks when
    the need falls below half of it.  Identical on all ranks for the pair capacity (every rank sees the same `need`).  Stable
    sizes let the caching allocator hand back the same blocks every frame (a capacity that tracks the need closely changes
    every frame and costs a cudaMalloc + implicit synchronisation each time)."""
    cur = st.get(key) or 0
    if cur < need + 256 or cur > 2 * (need + 256) + 8192:
        cur = (int(need * 1.25) + 256 + 4095) // 4096 * 4096
    cur = min(cur, limit) if limit >= need else cur
    st[key] = cur
    return cur


class _ShardedRasterizeNative(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, rs, group, deferred):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = means3D.device
        H, W = int(rs.image_height), int(rs.image_width)
        tiles_y, maxh = _strip_rows(H, world)
        if world > tiles_y:
            raise ValueError(f"{world} ranks but only {tiles_y} tile rows: a rank would own an empty strip")
        P = int(means3D.shape[0])
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        means3D = _dev_f32(means3D, dev)
        opacities = _dev_f32(opacities, dev)
        sh = _dev_f32(sh, dev) if sh.numel() else sh
        colors_precomp = _dev_f32(colors_precomp, dev) if colors_precomp.numel() else colors_precomp
        scales = _dev_f32(scales, dev) if scales.numel() else scales
        rotations = _dev_f32(rotations, dev) if rotations.numel() else rotations
        cov3Ds_precomp = _dev_f32(cov3Ds_precomp, dev) if cov3Ds_precomp.numel() else cov3Ds_precomp
        M = int(sh.size(1)) if sh.numel() else 0
        st = _state((str(dev), W, H, world))
        comm = native_comm(group, dev)
        if st["P"] != P or st["Pmax"] is None:
            sizes = torch.tensor([P], dtype=torch.int64, device=dev)
            all_sizes = torch.empty((world,), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_sizes, sizes, group=group)
            st["Pmax"], st["P"] = max(1, int(all_sizes.max().item())), P
        Pmax = st["Pmax"]
        rb, re_ = strip_bounds(tiles_y, world, rank)
        keep = []
        with torch.cuda.device(dev):
            nctx = _context(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            full = _make_frame(rs, dev, M, keep)
            strip = _make_frame(rs, dev, 0, keep)
            strip.tile_rank, strip.tile_world, strip.tile_row_begin, strip.tile_row_end = 0, 1, rb, re_
            g = _make_gaussians(P, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp)
            Pp = max(P, 1)
            geom_local = torch.empty((_lib.g4r_geom_bytes(Pp),), **u8)
            local_i = torch.zeros((2, Pp), **i32) if P == 0 else torch.empty((2, Pp), **i32)        # radii | n_touched
            pack_scratch = torch.empty((_lib.g4r_shard_scratch_bytes(Pp, world) + 64,), **u8)
            images = torch.empty((PLANES, H, W), **f32)
            img_state = torch.empty((_lib.g4r_image_bytes(W, H),), **u8)
            strip_elems = PLANES * maxh * W
            for _attempt in range(3):
                cap = _sticky(st, "cap_alloc", st["pair_hint"], Pmax) if st["pair_hint"] > 0 else Pmax
                rows = cap + 1
                counts_offset = strip_elems + world * rows
                payload_elems = (counts_offset + world + 3) // 4 * 4
                payload = torch.empty((payload_elems,), **f32)
                slabs = torch.empty((2, world, rows, REC_FLOATS), **f32)                             # send | recv
                slots = torch.empty((world, Pp), **i32)
                gathered = torch.empty((world * rows,), **f32)                                       # n_touched segments coming back
                cap_n = _sticky(st, "cap_n_alloc", st["n_hint"], 1 << 40) if st["n_hint"] > 0 else max(4096, 6 * cap)
                binning = torch.empty((_lib.g4r_binning_bytes(cap_n),), **u8)
                sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap_n),), **u8)
                all_i = torch.empty((world * rows + 4,), **i32)                                      # radii_all | worst
                b = _ShardBuffers(world, rank, cap, cap_n, geom_local.data_ptr(), local_i.data_ptr(), local_i.data_ptr() + 4 * Pp,
                                  slabs.data_ptr(), slabs.data_ptr() + 4 * world * rows * REC_FLOATS, slots.data_ptr(), pack_scratch.data_ptr(),
                                  payload.data_ptr(), strip_elems, maxh, counts_offset, payload_elems, gathered.data_ptr(), images.data_ptr(),
                                  img_state.data_ptr(), binning.data_ptr(), sort_scratch.data_ptr(), all_i.data_ptr(),
                                  all_i.data_ptr() + 4 * world * rows)
                _settle(st, nctx)           # a deferred check of the previous frame (raises if that frame had overflowed)
                _check(_lib.g4r_shard_forward_a(comm, nctx, ctypes.byref(full), ctypes.byref(strip), ctypes.byref(g), ctypes.byref(b), stream))
                if deferred and st["pair_hint"] > 0 and st["n_hint"] > 0:
                    # Deferred check: nothing is read back now.  The two early events of this frame are examined when the next
                    # forward (or this frame's backward) starts -- by then they have long fired -- so the host never waits on the
                    # device inside a frame and runs a whole frame ahead.  Capacities carry >= 25 % headroom with hysteresis.
                    st["pending"] = (cap, cap_n)
                    _check(_lib.g4r_shard_forward_b(comm, ctypes.byref(full), P, ctypes.byref(b), stream))
                    break
                N, worst = ctypes.c_int64(), ctypes.c_int64()
                _check(_lib.g4r_shard_forward_wait(nctx, ctypes.byref(N), ctypes.byref(worst)))
                N, worst = int(N.value), int(worst.value)
                st["pair_hint"] = max(worst, int(st["pair_hint"] * 0.95))
                st["n_hint"] = max(N, int(st["n_hint"] * 0.95))
                st["cap"] = cap
                if worst > cap:
                    st["redos"] += 1             # some pair's slab was too small: every rank sees the same `worst` and redoes part a
                    continue
                if N > cap_n:                    # local: only this rank's strip outgrew its instance buffers
                    cap_n = N
                    binning = torch.empty((_lib.g4r_binning_bytes(cap_n),), **u8)
                    sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap_n),), **u8)
                    b.cap_n, b.binning, b.sort_scratch = cap_n, binning.data_ptr(), sort_scratch.data_ptr()
                    _check(_lib.g4r_shard_render_owned(nctx, ctypes.byref(strip), ctypes.byref(b), stream))
                _check(_lib.g4r_shard_forward_b(comm, ctypes.byref(full), P, ctypes.byref(b), stream))
                break
            else:
                raise RuntimeError("sharded render: slab capacity kept overflowing")
        ctx.settle = (st, nctx)
        ctx.rs, ctx.group, ctx.P, ctx.M, ctx.world, ctx.rank, ctx.cap, ctx.cap_n = rs, group, P, M, world, rank, cap, cap_n
        ctx.frames = (full, strip, keep)
        ctx.opacities_shape = tuple(opacities.shape)
        ctx.geo = (strip_elems, maxh, counts_offset, payload_elems)
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, local_i, geom_local, slabs, img_state, slots, binning)
        radii = local_i[0, :P]
        n_touched = local_i[1, :P]
        ctx.mark_non_differentiable(radii, n_touched)
        return images[0:3], radii, images[3:4], images[4:5], n_touched

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_opacity, grad_ntouched):
        means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, local_i, geom_local, slabs, img_state, slots, binning = ctx.saved_tensors
        group, P, M, world, rank, cap = ctx.group, ctx.P, ctx.M, ctx.world, ctx.rank, ctx.cap
        _settle(*ctx.settle)             # deferred mode: the forward's capacity check, now that its events have fired
        full, strip, _keep = ctx.frames
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        grad_color = _dev_f32(grad_color, dev)
        grad_depth = _dev_f32(grad_depth, dev)
        rows = cap + 1
        Pp = max(P, 1)
        acc = torch.empty((2 * world * rows + Pp, ACC_FLOATS), **f32)                                 # acc_all | acc_back | acc_local
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
        tau = torch.empty((8,), **f32)
        needs = ctx.needs_input_grad
        with torch.cuda.device(dev):
            comm = native_comm(group, dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            g = _make_gaussians(P, means3D, means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp)
            b = _ShardBuffers(world, rank, cap, ctx.cap_n, geom_local.data_ptr(), local_i.data_ptr(), local_i.data_ptr() + 4 * Pp, None,
                              slabs.data_ptr() + 4 * world * rows * REC_FLOATS, slots.data_ptr(), None, None, ctx.geo[0], ctx.geo[1], ctx.geo[2],
                              ctx.geo[3], None, None, img_state.data_ptr(), binning.data_ptr(), None, None, None)
            io = _BackwardIO(None, None, _ptr(grads["means3D"]), _ptr(grads["means2D"]), _ptr(grads["opacities"]), _ptr(grads.get("sh")),
                             _ptr(grads.get("colors")), _ptr(grads.get("scales")), _ptr(grads.get("rots")), _ptr(grads.get("cov")), tau.data_ptr())
            a0 = acc.data_ptr()
            step = 4 * ACC_FLOATS * world * rows
            _check(_lib.g4r_shard_backward(comm, ctypes.byref(full), ctypes.byref(strip), ctypes.byref(g), ctypes.byref(b), grad_color.data_ptr(),
                                           grad_depth.data_ptr(), a0, a0 + step, a0 + 2 * step, ctypes.byref(io), 1 if (needs[8] or needs[9]) else 0,
                                           stream))
        return (grads["means3D"], grads["means2D"], grads.get("sh"), grads.get("colors"), grads["opacities"], grads.get("scales"),
                grads.get("rots"), grads.get("cov"), tau[3:6].view(1, -1) if needs[8] else None, tau[:3].view(1, -1) if needs[9] else None,
                None, None, None)


class ShardedGaussianRasterizer(torch.nn.Module):
    """Same call signature as ``GaussianRasterizer`` but every argument is the LOCAL shard of the Gaussians; returns the
    full image on every rank and the local ``radii`` / ``n_touched``.  The image gradients handed to backward must be
    identical on all ranks (every rank evaluates the loss on the full image)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings, group=None, backend=None, exchange: str = "alltoall", native: bool = True,
                 deferred_check: bool = False):
        """`backend` = None: the CUDA library.  `native` (only without an injected backend): True = the frame is enqueued by the
        native runtime of csrc/shard_nccl.cu (own NCCL communicator, 4 C calls per fwd+bwd); False = the same steps orchestrated
        from Python through torch.distributed (the path the gloo tests exercise with a CPU backend)."""
        super().__init__()
        if exchange != "alltoall":
            raise ValueError("exchange must be 'alltoall' (the all-gather variant of round 1 moved world x more bytes and was removed)")
        self.raster_settings = raster_settings
        self.group = group
        self.native = bool(native) and backend is None and os.environ.get("G4R_SHARD_NATIVE", "1") != "0"      # env: A/B switch
        # deferred_check (native runtime only): the capacity checks of a frame are made when the NEXT frame (or this frame's
        # backward) starts instead of in the middle of the forward, so the host never waits for the device inside a frame.
        # A frame that outgrew its buffers then raises one call late (its outputs were invalid) -- for loops whose instance
        # counts drift slowly (SLAM mapping) and that can re-render a frame.
        self.deferred = bool(deferred_check) or os.environ.get("G4R_SHARD_DEFERRED", "0") == "1"
        self.backend = backend if backend is not None else CudaBackend()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                theta=None, rho=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        e = lambda t: torch.Tensor([]) if t is None else t
        if self.native:
            return _ShardedRasterizeNative.apply(means3D, means2D, e(shs), e(colors_precomp), opacities, e(scales), e(rotations), e(cov3D_precomp),
                                                 e(theta), e(rho), self.raster_settings, self.group, self.deferred)
        return _ShardedRasterize.apply(means3D, means2D, e(shs), e(colors_precomp), opacities, e(scales), e(rotations), e(cov3D_precomp),
                                       e(theta), e(rho), self.raster_settings, self.group, self.backend)
