"""B200-native drop-in for the ``diff_gaussian_rasterization`` extension of yanyan-li/4DGS-SLAM.

Python surface mirrors the reference package line for line in *behaviour*
(``/root/reference/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py``):

* ``GaussianRasterizationSettings``  -- reference ``__init__.py:173-186`` (same 13 fields, same order)
* ``GaussianRasterizer``             -- reference ``__init__.py:188-244`` (``forward`` / ``markVisible``)
* ``rasterize_gaussians``            -- reference ``__init__.py:21-46``
* ``_RasterizeGaussians``            -- reference ``__init__.py:48-171`` (autograd op, same input / gradient order)

so ``gaussian_splatting/gaussian_renderer/__init__.py:15-18,87-106,180-214`` (``render`` / ``render_flow``),
``utils/slam_frontend.py`` and ``utils/slam_backend.py`` run unchanged on top of it.

The native side is NOT a torch extension: it is a plain C-ABI shared library (``libg4r.so``, built from
``4dgs-slam_b200/csrc/*.cu`` for sm_100a, declared in ``include/g4r.h``) loaded with ctypes.  torch only
provides device memory and the current stream.  There is no CPU or PyTorch fallback: if the library is
missing, importing this package raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

__all__ = [
    "GaussianRasterizationSettings",
    "GaussianRasterizer",
    "rasterize_gaussians",
    "rasterize_gaussians_with_state",
    "FusedGaussianRasterizer",
    "rasterize_gaussians_raw",
    "dynamic_slots",
    "captured_overflow",
    "reset_captured",
    "host_backend",
]

# ----------------------------------------------------------------------------------------------
# native library
# ----------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libg4r.so")


class _Frame(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float), ("scale_modifier", ctypes.c_float),
        ("sh_degree", ctypes.c_int32), ("sh_coeffs", ctypes.c_int32), ("prefiltered", ctypes.c_int32),
        ("bg", ctypes.c_void_p), ("viewmatrix", ctypes.c_void_p), ("projmatrix", ctypes.c_void_p),
        ("projmatrix_raw", ctypes.c_void_p), ("campos", ctypes.c_void_p),
        ("tile_rank", ctypes.c_int32), ("tile_world", ctypes.c_int32),
        ("tile_row_begin", ctypes.c_int32), ("tile_row_end", ctypes.c_int32),
    ]


class _Gaussians(ctypes.Structure):
    _fields_ = [
        ("P", ctypes.c_int32),
        ("means3D", ctypes.c_void_p), ("opacities", ctypes.c_void_p), ("shs", ctypes.c_void_p),
        ("colors_precomp", ctypes.c_void_p), ("scales", ctypes.c_void_p), ("rotations", ctypes.c_void_p),
        ("cov3D_precomp", ctypes.c_void_p),
        ("shs_rest", ctypes.c_void_p), ("activation", ctypes.c_int32), ("scale_dim", ctypes.c_int32),
        ("mask", ctypes.c_void_p), ("dyn_slot", ctypes.c_void_p), ("dx", ctypes.c_void_p), ("ds", ctypes.c_void_p), ("dr", ctypes.c_void_p),
    ]


class _ForwardOut(ctypes.Structure):
    _fields_ = [
        ("color", ctypes.c_void_p), ("depth", ctypes.c_void_p), ("opacity", ctypes.c_void_p),
        ("radii", ctypes.c_void_p), ("n_touched", ctypes.c_void_p), ("color_plane_stride", ctypes.c_int64),
    ]


class _BackwardIO(ctypes.Structure):
    _fields_ = [
        ("dL_dcolor", ctypes.c_void_p), ("dL_ddepth", ctypes.c_void_p),
        ("dL_dmeans3D", ctypes.c_void_p), ("dL_dmeans2D", ctypes.c_void_p), ("dL_dopacity", ctypes.c_void_p),
        ("dL_dshs", ctypes.c_void_p), ("dL_dcolors_precomp", ctypes.c_void_p), ("dL_dscales", ctypes.c_void_p),
        ("dL_drotations", ctypes.c_void_p), ("dL_dcov3D", ctypes.c_void_p), ("dL_dtau", ctypes.c_void_p),
        ("dL_dshs_rest", ctypes.c_void_p), ("dL_ddx", ctypes.c_void_p), ("dL_dds", ctypes.c_void_p), ("dL_ddr", ctypes.c_void_p),
    ]


class _Layout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in (
        "geom_rec", "geom_clamped", "img_final_T", "img_n_contrib", "img_ranges", "img_counts", "img_header",
        "bin_point_list", "bin_pairs")]


def _load_library() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"diff_gaussian_rasterization (B200-native): {_LIB_PATH} is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` from the repository root "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU / PyTorch fallback."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
    lib.g4r_last_error.restype = ctypes.c_char_p
    lib.g4r_last_error.argtypes = []
    lib.g4r_version.restype = ctypes.c_int
    lib.g4r_version.argtypes = []
    lib.g4r_context_create.restype = ctypes.c_int
    lib.g4r_context_create.argtypes = [ctypes.POINTER(vp)]
    lib.g4r_context_destroy.restype = None
    lib.g4r_context_destroy.argtypes = [vp]
    lib.g4r_geom_bytes.restype = sz
    lib.g4r_geom_bytes.argtypes = [i32]
    lib.g4r_image_bytes.restype = sz
    lib.g4r_image_bytes.argtypes = [i32, i32]
    lib.g4r_binning_bytes.restype = sz
    lib.g4r_binning_bytes.argtypes = [i64]
    lib.g4r_sort_scratch_bytes.restype = sz
    lib.g4r_sort_scratch_bytes.argtypes = [i64]
    lib.g4r_backward_scratch_bytes.restype = sz
    lib.g4r_backward_scratch_bytes.argtypes = [i32]
    lib.g4r_forward_project.restype = ctypes.c_int
    lib.g4r_forward_project.argtypes = [vp, ctypes.POINTER(_Frame), ctypes.POINTER(_Gaussians), vp, vp, vp, vp, vp]
    lib.g4r_forward_render.restype = ctypes.c_int
    lib.g4r_forward_render.argtypes = [vp, ctypes.POINTER(_Frame), ctypes.POINTER(_Gaussians), vp, vp, vp, vp, i64,
                                       ctypes.POINTER(_ForwardOut), vp]
    lib.g4r_wait_num_rendered.restype = i64
    lib.g4r_wait_num_rendered.argtypes = [vp]
    lib.g4r_backward.restype = ctypes.c_int
    lib.g4r_backward.argtypes = [ctypes.POINTER(_Frame), ctypes.POINTER(_Gaussians), vp, vp, vp, vp, vp,
                                 ctypes.POINTER(_BackwardIO), vp]
    lib.g4r_mark_visible.restype = ctypes.c_int
    lib.g4r_mark_visible.argtypes = [i32, vp, vp, vp, vp, vp]
    lib.g4r_overflow_status.restype = i64
    lib.g4r_overflow_status.argtypes = [ctypes.c_int]
    lib.g4r_layout.restype = ctypes.c_int
    lib.g4r_layout.argtypes = [i32, i32, i32, i64, ctypes.POINTER(_Layout)]
    # the ctypes mirrors above must have exactly the C layout of include/g4r.h
    sizes = (ctypes.c_int32 * 5)()
    lib.g4r_struct_sizes(sizes)
    mine = [ctypes.sizeof(c) for c in (_Frame, _Gaussians, _ForwardOut, _BackwardIO, _Layout)]
    if list(sizes) != mine:
        raise ImportError(f"diff_gaussian_rasterization: struct layout mismatch between libg4r.so {list(sizes)} and the Python "
                          f"bindings {mine}; rebuild the library")
    if lib.g4r_version() != 5:
        raise ImportError(f"diff_gaussian_rasterization: libg4r.so has ABI version {lib.g4r_version()}, expected 5; rebuild it")
    return lib


_lib = _load_library()
LIBRARY_PATH = _LIB_PATH


def _check(rc: int) -> None:
    if rc < 0:
        raise RuntimeError(_lib.g4r_last_error().decode("utf-8", "replace"))


_tls = threading.local()


def _context(device: torch.device) -> int:
    """One native context (pinned int + event) per host thread and device."""
    table = getattr(_tls, "ctx", None)
    if table is None:
        table = _tls.ctx = {}
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = table.get(idx)
    if h is None:
        out = ctypes.c_void_p()
        _check(_lib.g4r_context_create(ctypes.byref(out)))
        h = table[idx] = out.value
    return h


# instance-capacity hint per (device, W, H): high-water mark of num_rendered with slow decay.  The autograd engine runs
# backward on its own threads and applications may render from several threads, hence the lock.
_state_lock = threading.Lock()
_cap_hint: dict = {}

# CUDA-graph capture: no host read-back is possible, so phase 2 gets capacity = hint x this factor.
_GRAPH_CAPACITY_FACTOR = 2.0


def captured_overflow(device=None) -> bool:
    """True if, since the last call, a forward that ran WITHOUT host read-back (i.e. captured in a CUDA graph and replayed)
    produced more (tile, Gaussian) instances than the capacity fixed at capture time.  Such a replay leaves the forward
    outputs stale and its backward returns zero gradients (the kernels compare N with the capacity on the device and
    exit untouched).  The flag lives in the library (one word per device, set by the kernel that detects the overflow), so
    it does not depend on which graphs are still alive.  Synchronises the device and clears the flag; after a True, run a
    few eager iterations (they update the capacity hint) and capture again."""
    devs = [torch.device(device).index] if device is not None else list(range(torch.cuda.device_count()))
    bad = False
    for d in devs:
        with torch.cuda.device(d if d is not None else torch.cuda.current_device()):
            n = int(_lib.g4r_overflow_status(1))
            if n < 0:
                _check(n)
            bad = bad or n > 0
    return bad


def reset_captured() -> None:
    """Clears the overflow flag on every device (kept for API compatibility with round 1)."""
    captured_overflow()


class _on_device:
    """`with torch.cuda.device(d)` costs two cudaSetDevice round trips (~10 us); nearly every call already runs on the tensors'
    device, so the guard is only entered when it is needed."""
    __slots__ = ("guard",)

    def __init__(self, device: torch.device):
        idx = device.index
        self.guard = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


_size_cache: dict = {}


def _bytes(fn_name: str, *args) -> int:
    """Scratch sizes from the library, memoised (they are pure functions of their arguments)."""
    key = (fn_name,) + args
    v = _size_cache.get(key)
    if v is None:
        if len(_size_cache) > 4096:
            _size_cache.clear()
        v = _size_cache[key] = int(getattr(_lib, fn_name)(*args))
    return v


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _dev_f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    if t.dtype is torch.float32 and t.device == device:
        return t if t.is_contiguous() else t.contiguous()
    return t.to(device=device, dtype=torch.float32).contiguous()


def _make_frame(rs: "GaussianRasterizationSettings", device: torch.device, sh_coeffs: int, keep: list) -> _Frame:
    def mat(t):
        t = _dev_f32(t, device)
        keep.append(t)
        return t.data_ptr()

    f = _Frame()
    f.width, f.height = int(rs.image_width), int(rs.image_height)
    f.tan_fovx, f.tan_fovy = float(rs.tanfovx), float(rs.tanfovy)
    f.scale_modifier = float(rs.scale_modifier)
    f.sh_degree, f.sh_coeffs = int(rs.sh_degree), int(sh_coeffs)
    f.prefiltered = int(bool(rs.prefiltered))
    f.bg, f.viewmatrix, f.projmatrix = mat(rs.bg), mat(rs.viewmatrix), mat(rs.projmatrix)
    f.projmatrix_raw, f.campos = mat(rs.projmatrix_raw), mat(rs.campos)
    f.tile_rank, f.tile_world = 0, 1
    f.tile_row_begin, f.tile_row_end = 0, 0
    return f


def _make_gaussians(P, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, raw=None, extra=None) -> _Gaussians:
    """`raw` = None (reference surface: activated parameters) or (features_rest | None, scale_dim): include/g4r.h G4R_ACT_RAW.
    `extra` = None or dict(mask=uint8 [P] | None, dyn_slot=int32 [P] | None, dx, ds, dr): the in-kernel static mask / dynamic offsets."""
    g = _Gaussians()
    if extra is not None:
        g.mask, g.dyn_slot = _ptr(extra.get("mask")), _ptr(extra.get("dyn_slot"))
        g.dx, g.ds, g.dr = _ptr(extra.get("dx")), _ptr(extra.get("ds")), _ptr(extra.get("dr"))
    g.P = P
    g.means3D, g.opacities = _ptr(means3D), _ptr(opacities)
    g.shs, g.colors_precomp = _ptr(sh), _ptr(colors_precomp)
    g.scales, g.rotations, g.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp)
    if raw is not None:
        g.activation = 1
        g.shs_rest = _ptr(raw[0])
        g.scale_dim = int(raw[1])
    return g


def _layout(P: int, W: int, H: int, cap: int) -> _Layout:
    lay = _Layout()
    _check(_lib.g4r_layout(P, W, H, cap, ctypes.byref(lay)))
    return lay


# ----------------------------------------------------------------------------------------------
# forward / backward drivers
# ----------------------------------------------------------------------------------------------
def _forward_impl(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, raw=None, extra=None):
    """`raw` = None, or (features_rest tensor | None, scale_dim) when the tensors are GaussianModel's raw parameters
    (then `sh` is _features_dc [P,1,3]).  `extra` = None or the static mask / dynamic offsets (see _make_gaussians)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")   # rasterize_points.cu:58-60
    if not means3D.is_cuda:
        raise RuntimeError("diff_gaussian_rasterization (B200-native): means3D must be a CUDA tensor; there is no CPU path")
    device = means3D.device
    P = int(means3D.size(0))
    H, W = int(rs.image_height), int(rs.image_width)

    means3D = _dev_f32(means3D, device)
    opacities = _dev_f32(opacities, device)
    sh = _dev_f32(sh, device) if sh.numel() else sh
    colors_precomp = _dev_f32(colors_precomp, device) if colors_precomp.numel() else colors_precomp
    scales = _dev_f32(scales, device) if scales.numel() else scales
    rotations = _dev_f32(rotations, device) if rotations.numel() else rotations
    cov3Ds_precomp = _dev_f32(cov3Ds_precomp, device) if cov3Ds_precomp.numel() else cov3Ds_precomp
    M = int(sh.size(1)) if sh.numel() else 0                                    # rasterize_points.cu:87-91
    if raw is not None:
        rest = raw[0]
        if rest is not None and rest.numel():
            rest = _dev_f32(rest, device)
            M = 1 + int(rest.size(1))
        else:
            rest = None
        raw = (rest, raw[1])

    f32 = dict(dtype=torch.float32, device=device)
    i32 = dict(dtype=torch.int32, device=device)
    u8 = dict(dtype=torch.uint8, device=device)
    if P == 0:
        # rasterize_points.cu:69-73,85: zero images, no kernels
        z = torch.zeros
        state = dict(P=0, N=0, geom=torch.empty(0, **u8), img=torch.empty(0, **u8), binning=torch.empty(0, **u8), capacity=0)
        return z((3, H, W), **f32), torch.zeros((0,), **i32), z((1, H, W), **f32), z((1, H, W), **f32), torch.zeros((0,), **i32), state

    # Phase 1 needs only the per-Gaussian state: it is enqueued first, and the host allocates everything phase 2 writes while
    # the projection kernel already runs (the GPU is idle at this point whenever the caller is host-bound, e.g. during tracking).
    ints = torch.empty((2, P), **i32)                       # radii | n_touched
    radii, n_touched = ints[0], ints[1]
    geom = torch.empty((_bytes("g4r_geom_bytes", P),), **u8)
    img = torch.empty((_bytes("g4r_image_bytes", W, H),), **u8)

    keep: list = []
    with _on_device(device):
        capturing = torch.cuda.is_current_stream_capturing()
        ctx = None if capturing else _context(device)     # (creating one allocates pinned memory: not allowed while capturing)
        stream = torch.cuda.current_stream(device).cuda_stream
        frame = _make_frame(rs, device, M, keep)
        g = _make_gaussians(P, means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, raw, extra)
        _check(_lib.g4r_forward_project(ctx, ctypes.byref(frame), ctypes.byref(g), geom.data_ptr(), img.data_ptr(),
                                        radii.data_ptr(), n_touched.data_ptr(), stream))
        # three separate tensors like the reference's: views of one buffer would make an in-place operation on any of them an
        # autograd error ("a view ... of a function that returns multiple views")
        color = torch.empty((3, H, W), **f32)
        depth = torch.empty((1, H, W), **f32)
        opacity = torch.empty((1, H, W), **f32)
        out = _ForwardOut(color.data_ptr(), depth.data_ptr(), opacity.data_ptr(), radii.data_ptr(), n_touched.data_ptr())

        key = (device.index, W, H)
        with _state_lock:
            hint = _cap_hint.get(key, 0)
        if capturing:
            # CUDA-graph capture (torch.cuda.graph around forward + loss + backward): no host read-back is possible, so
            # phase 2 gets a generous fixed capacity derived from the eager warm-up iterations.  A replay whose instance
            # count outgrows it leaves the outputs untouched, its backward returns zeros, and `captured_overflow()` reports it.
            cap = int(max(hint, 4 * P) * _GRAPH_CAPACITY_FACTOR) + 4096
            binning = torch.empty((_lib.g4r_binning_bytes(cap),), **u8)
            sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap),), **u8)
            _check(_lib.g4r_forward_render(None, ctypes.byref(frame), ctypes.byref(g), geom.data_ptr(), img.data_ptr(),
                                           binning.data_ptr(), sort_scratch.data_ptr(), cap, ctypes.byref(out), stream))
            if raw is not None and raw[0] is not None:
                keep.append(raw[0])
            state = dict(P=P, N=-1, geom=geom, img=img, binning=binning, capacity=cap, frame=(frame, keep))
            return color, radii, depth, opacity, n_touched, state

        # Phase 2 is enqueued with a speculative capacity; N is checked only after everything is in the
        # stream, so the device never waits for the host (the reference blocks on a cudaMemcpy instead,
        # rasterizer_impl.cu:284).  `binning` (the sorted id list) is saved for backward; `sort_scratch` (the unsorted
        # pairs, 4x larger) dies with this call -- the caching allocator recycles it stream-ordered.
        cap = int(max(hint, 4 * P) * 1.25) + 4096 if hint == 0 else int(hint * 1.25) + 4096
        binning = torch.empty((_lib.g4r_binning_bytes(cap),), **u8)
        sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap),), **u8)
        _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), geom.data_ptr(), img.data_ptr(),
                                       binning.data_ptr(), sort_scratch.data_ptr(), cap, ctypes.byref(out), stream))
        N = int(_lib.g4r_wait_num_rendered(ctx))
        if N < 0:
            _check(N)
        if N > cap:
            cap = N
            binning = torch.empty((_lib.g4r_binning_bytes(cap),), **u8)
            sort_scratch = torch.empty((_lib.g4r_sort_scratch_bytes(cap),), **u8)
            _check(_lib.g4r_forward_render(ctx, ctypes.byref(frame), ctypes.byref(g), geom.data_ptr(), img.data_ptr(),
                                           binning.data_ptr(), sort_scratch.data_ptr(), cap, ctypes.byref(out), stream))
        with _state_lock:
            _cap_hint[key] = max(N, int(_cap_hint.get(key, 0) * 0.95))
    if raw is not None and raw[0] is not None:
        keep.append(raw[0])
    state = dict(P=P, N=N, geom=geom, img=img, binning=binning, sort_scratch=sort_scratch, capacity=cap, frame=(frame, keep))
    return color, radii, depth, opacity, n_touched, state


def _backward_impl(rs, P, means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii, geom, img, binning,
                   opacities_shape, grad_out_color, grad_out_depth, frame_keep=None, raw=None, want=None, extra=None):
    """`raw` = None or (features_rest | None, scale_dim): gradients are then w.r.t. the raw parameters and a tenth element, the
    gradient of features_rest, is appended to the returned tuple.  `want` = dict of booleans (means3D, means2D, sh,
    colors, opacities, scales, rotations, cov) from autograd's needs_input_grad: gradients nobody asked for are neither
    allocated nor written (pose tracking only consumes dL/dtau, utils/slam_frontend.py:441-448); missing keys mean True."""
    device = means3D.device
    f32 = dict(dtype=torch.float32, device=device)
    w = (lambda k: True) if want is None else (lambda k: bool(want.get(k, True)))
    M = int(sh.size(1)) if sh.numel() else 0
    grad_rest = None
    if raw is not None and raw[0] is not None:
        M = 1 + int(raw[0].size(1))
        if w("sh"):
            grad_rest = torch.empty((P, M - 1, 3), **f32)
    tau = torch.empty((8,), **f32)
    grad_means3D = torch.empty((P, 3), **f32) if w("means3D") else None
    grad_means2D = torch.empty((P, 3), **f32) if w("means2D") else None
    grad_opacities = torch.empty(opacities_shape, **f32) if w("opacities") else None
    grad_sh = torch.empty((P, 1 if raw is not None else M, 3), **f32) if sh.numel() and w("sh") else None
    grad_colors = torch.empty((P, 3), **f32) if colors_precomp.numel() and w("colors") else None
    grad_scales = torch.empty((P, raw[1] if raw is not None else 3), **f32) if scales.numel() and w("scales") else None
    grad_rot = torch.empty((P, 4), **f32) if rotations.numel() and w("rotations") else None
    grad_cov = torch.empty((P, 6), **f32) if cov3Ds_precomp.numel() and w("cov") else None
    grad_off = [None, None, None]
    if extra is not None and extra.get("dyn_slot") is not None:
        for k, name in enumerate(("dx", "ds", "dr")):
            if extra.get(name) is not None and w(name):
                grad_off[k] = torch.empty_like(extra[name])
    if P == 0:
        tau.zero_()
        res = (grad_means3D, grad_means2D, grad_sh, grad_colors, grad_opacities, grad_scales, grad_rot, grad_cov, tau)
        if extra is not None:
            return res + (grad_rest, [None if g_ is None else g_.zero_() for g_ in grad_off])
        return res + (grad_rest,) if raw is not None else res

    grad_out_color = _dev_f32(grad_out_color, device)
    grad_out_depth = _dev_f32(grad_out_depth, device)
    scratch = torch.empty((_bytes("g4r_backward_scratch_bytes", P),), dtype=torch.uint8, device=device)
    with _on_device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        # the forward's frame struct (camera pointers; the tensors behind them are kept alive next to it) is reused
        frame, keep = frame_keep if frame_keep is not None else (None, [])
        if frame is None:
            frame = _make_frame(rs, device, M, keep)
        g = _make_gaussians(P, means3D, None, sh, colors_precomp, scales, rotations, cov3Ds_precomp,
                            None if raw is None else (raw[0], raw[1]), extra)
        g.opacities = means3D.data_ptr()   # opacities are not read in backward (they live in the splat records)
        io = _BackwardIO(grad_out_color.data_ptr(), grad_out_depth.data_ptr(), _ptr(grad_means3D), _ptr(grad_means2D),
                         _ptr(grad_opacities), _ptr(grad_sh), _ptr(grad_colors), _ptr(grad_scales), _ptr(grad_rot),
                         _ptr(grad_cov), tau.data_ptr(), _ptr(grad_rest), _ptr(grad_off[0]), _ptr(grad_off[1]), _ptr(grad_off[2]))
        _check(_lib.g4r_backward(ctypes.byref(frame), ctypes.byref(g), radii.data_ptr(), geom.data_ptr(), img.data_ptr(),
                                 binning.data_ptr(), scratch.data_ptr(), ctypes.byref(io), stream))
    res = (grad_means3D, grad_means2D, grad_sh, grad_colors, grad_opacities, grad_scales, grad_rot, grad_cov, tau)
    if extra is not None:
        return res + (grad_rest, grad_off)
    return res + (grad_rest,) if raw is not None else res


# ----------------------------------------------------------------------------------------------
# public surface (same names / order as the reference package)
# ----------------------------------------------------------------------------------------------
# The same op with its host side in C++ (csrc/host/g4r_torch.cpp -> _g4r_host.so, built by __graft_entry__.build_host): a
# torch::autograd::Function over the same C ABI, ~0.15 ms less host time per fwd+bwd.  Optional: without it (or with
# G4R_HOST=python) everything runs through the Python Function below, which stays the specification; debug=True, CUDA-graph
# capture and every opt-in extension always do.
_host = None
if os.environ.get("G4R_HOST", "cpp").lower() != "python":
    try:
        from . import _g4r_host as _host
        _sizes = (ctypes.c_int32 * 5)()
        _lib.g4r_struct_sizes(_sizes)
        if _host.abi_version() != 5 or list(_host.struct_sizes()) != list(_sizes)[:4]:
            _host = None                       # built against another g4r.h: rebuild with __graft_entry__.build_host()
    except ImportError:
        _host = None


def host_backend() -> str:
    """"cpp" when the standard op dispatches to the C++ host extension, else "python"."""
    return "cpp" if _host is not None else "python"


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                        raster_settings):
    rs = raster_settings
    if (_host is not None and means3D.is_cuda and not getattr(rs, "debug", False) and not torch.cuda.is_current_stream_capturing()):
        key = (means3D.device.index, int(rs.image_width), int(rs.image_height))
        with _state_lock:
            hint = _cap_hint.get(key, 0)
        color, radii, depth, opacity, n_touched, N = _host.rasterize(
            means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, rs.bg, rs.viewmatrix, rs.projmatrix,
            rs.projmatrix_raw, rs.campos, int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier),
            int(rs.sh_degree), bool(rs.prefiltered), hint)
        if means3D.shape[0]:
            with _state_lock:
                _cap_hint[key] = max(N, int(_cap_hint.get(key, 0) * 0.95))
        return color, radii, depth, opacity, n_touched
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     theta, rho, raster_settings)


def _cpu_copy(args):
    """cpu_deep_copy_tuple of the reference (__init__.py:18-20): the tensors of an argument tuple copied to the host."""
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def _debug_call(rs, which: str, args, fn):
    """`debug=True`: like the reference (__init__.py:90-97,141-148), the arguments are copied to the host before the call, and if
    the call (including the synchronisation behind it) raises they are written to snapshot_fw.dump / snapshot_bw.dump."""
    if not getattr(rs, "debug", False) or (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
        return fn()
    cpu_args = _cpu_copy(args)
    try:
        out = fn()
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                torch.cuda.synchronize(a.device)
                break
        return out
    except Exception:
        torch.save(cpu_args, f"snapshot_{which}.dump")
        print(f"\nAn error occured in {'forward' if which == 'fw' else 'backward'}. Writing snapshot_{which}.dump for debugging.\n")
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                raster_settings):
        color, radii, depth, opacity, n_touched, state = _debug_call(
            raster_settings, "fw", (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings),
            lambda: _forward_impl(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings))
        ctx.raster_settings = raster_settings
        ctx.num_rendered = state["N"]
        ctx.P = state["P"]
        ctx.opacities_shape = tuple(opacities.shape)
        ctx.frame_keep = state.get("frame")
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              state["geom"], state["binning"], state["img"])
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, opacity, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_radii, grad_out_depth, grad_out_opacity, grad_n_touched):
        # like the reference, grad_out_opacity / radii / n_touched are dropped (reference __init__.py:108-130)
        rs = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img = ctx.saved_tensors
        device = means3D.device
        means3D_c = _dev_f32(means3D, device)
        needs = ctx.needs_input_grad
        want = dict(means3D=needs[0], means2D=needs[1], sh=needs[2], colors=needs[3], opacities=needs[4], scales=needs[5],
                    rotations=needs[6], cov=needs[7])
        (grad_means3D, grad_means2D, grad_sh, grad_colors, grad_opacities, grad_scales, grad_rot, grad_cov, tau) = _debug_call(
            rs, "bw", (means3D, radii, colors_precomp, scales, rotations, cov3Ds_precomp, grad_out_color, grad_out_depth, sh, rs),
            lambda: _backward_impl(
                rs, ctx.P, means3D_c,
                _dev_f32(sh, device) if sh.numel() else sh,
                _dev_f32(colors_precomp, device) if colors_precomp.numel() else colors_precomp,
                _dev_f32(scales, device) if scales.numel() else scales,
                _dev_f32(rotations, device) if rotations.numel() else rotations,
                _dev_f32(cov3Ds_precomp, device) if cov3Ds_precomp.numel() else cov3Ds_precomp,
                radii, geom, img, binning, ctx.opacities_shape, grad_out_color, grad_out_depth, ctx.frame_keep, want=want))
        grad_rho = tau[:3].view(1, -1)
        grad_theta = tau[3:6].view(1, -1)
        return (
            grad_means3D,
            grad_means2D,
            grad_sh,
            grad_colors,
            grad_opacities,
            grad_scales,
            grad_rot,
            grad_cov,
            grad_theta if needs[8] else None,
            grad_rho if needs[9] else None,
            None,
        )


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


_EMPTY = torch.Tensor([])


def _empty() -> torch.Tensor:
    return _EMPTY


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points in front of the near plane (view-space z > 0.2)."""
        with torch.no_grad():
            rs = self.raster_settings
            if not positions.is_cuda:
                raise RuntimeError("diff_gaussian_rasterization (B200-native): positions must be a CUDA tensor")
            device = positions.device
            pos = _dev_f32(positions, device)
            P = int(pos.size(0))
            present = torch.zeros((P,), dtype=torch.bool, device=device)
            if P:
                view = _dev_f32(rs.viewmatrix, device)
                proj = _dev_f32(rs.projmatrix, device)
                with torch.cuda.device(device):
                    _check(_lib.g4r_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(), present.data_ptr(),
                                                 torch.cuda.current_stream(device).cuda_stream))
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None
        ):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")

        if shs is None:
            shs = _empty()
        if colors_precomp is None:
            colors_precomp = _empty()
        if scales is None:
            scales = _empty()
        if rotations is None:
            rotations = _empty()
        if cov3D_precomp is None:
            cov3D_precomp = _empty()
        if theta is None:
            theta = _empty()
        if rho is None:
            rho = _empty()

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   theta, rho, raster_settings)


# ----------------------------------------------------------------------------------------------
# opt-in extension: raw-parameter rasterizer ("render() prelude fusion")
# ----------------------------------------------------------------------------------------------
class _RasterizeGaussiansRaw(torch.autograd.Function):
    """Same op as `_RasterizeGaussians`, but the inputs are `GaussianModel`'s raw parameters
    (`_xyz, _features_dc, _features_rest, _opacity, _scaling, _rotation`); the activations of
    `gaussian_splatting/scene/gaussian_model.py:100-128` (sigmoid / exp / normalize / cat) run inside the projection kernel and
    their chain rule inside the per-Gaussian backward kernel, so the ~7 element-wise torch kernels of the prelude of
    `gaussian_renderer/__init__.py:108-131` and their autograd twins disappear.  Optionally the rest of that prelude too: the
    static `mask` (`:180-191`; here a per-Gaussian flag, no gather, outputs keep the full length) and the dynamic offsets
    `dx / ds / dr` (`:159-174`; here looked up through `dyn_slot`, no scatter into zero tensors)."""

    @staticmethod
    def forward(ctx, xyz, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, theta, rho, dx, ds, dr, raster_settings,
                mask, dyn_slot):
        e = _empty()
        scale_dim = int(scaling_raw.size(1)) if scaling_raw.dim() == 2 else 0
        if scale_dim not in (1, 3):
            raise RuntimeError("scaling_raw must have dimensions (num_points, 3) or (num_points, 1)")
        if features_dc.dim() != 3 or features_dc.size(1) != 1 or features_dc.size(2) != 3:
            raise RuntimeError("features_dc must have dimensions (num_points, 1, 3)")
        device = xyz.device
        extra = None
        if mask is not None or dyn_slot is not None:
            extra = dict(mask=None if mask is None else mask.to(device=device, dtype=torch.uint8).contiguous(),
                         dyn_slot=None if dyn_slot is None else dyn_slot.to(device=device, dtype=torch.int32).contiguous())
            if dyn_slot is not None:
                for name, t in (("dx", dx), ("ds", ds), ("dr", dr)):
                    extra[name] = _dev_f32(t, device) if t is not None and t.numel() else None
        color, radii, depth, opacity, n_touched, state = _forward_impl(
            xyz, features_dc, e, opacity_raw, scaling_raw, rotation_raw, e, raster_settings, raw=(features_rest, scale_dim), extra=extra)
        ctx.raster_settings = raster_settings
        ctx.P = state["P"]
        ctx.scale_dim = scale_dim
        ctx.opacities_shape = tuple(opacity_raw.shape)
        ctx.frame_keep = state.get("frame")
        ctx.extra = extra
        ctx.save_for_backward(xyz, features_dc, features_rest, scaling_raw, rotation_raw, radii,
                              state["geom"], state["binning"], state["img"])
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, opacity, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_radii, grad_out_depth, grad_out_opacity, grad_n_touched):
        xyz, features_dc, features_rest, scaling_raw, rotation_raw, radii, geom, binning, img = ctx.saved_tensors
        device = xyz.device
        e = _empty()
        rest = _dev_f32(features_rest, device) if features_rest.numel() else None
        needs = ctx.needs_input_grad
        want = dict(means3D=needs[0], means2D=needs[1], sh=needs[2] or needs[3], opacities=needs[4], scales=needs[5], rotations=needs[6],
                    dx=needs[9], ds=needs[10], dr=needs[11])
        res = _backward_impl(
            ctx.raster_settings, ctx.P, _dev_f32(xyz, device), _dev_f32(features_dc, device), e, _dev_f32(scaling_raw, device),
            _dev_f32(rotation_raw, device), e, radii, geom, img, binning, ctx.opacities_shape, grad_out_color, grad_out_depth,
            ctx.frame_keep, raw=(rest, ctx.scale_dim), want=want, extra=ctx.extra)
        (g_xyz, g_m2d, g_dc, _gc, g_op, g_sc, g_rot, _gcov, tau, g_rest) = res[:10]
        g_off = res[10] if len(res) > 10 else [None, None, None]
        if g_rest is None and features_rest.numel() == 0 and needs[3]:
            g_rest = torch.zeros_like(features_rest)
        return (g_xyz, g_m2d, g_dc, g_rest, g_op, g_sc, g_rot,
                tau[3:6].view(1, -1) if needs[7] else None, tau[:3].view(1, -1) if needs[8] else None,
                g_off[0], g_off[1], g_off[2], None, None, None)


def dynamic_slots(dygs: torch.Tensor) -> torch.Tensor:
    """int32 [P]: rank of every dynamic Gaussian among the dynamic ones (the row of dx / ds / dr that belongs to it), -1 for static
    Gaussians.  `dygs` only changes when the model is densified / pruned, so callers cache this next to it."""
    d = dygs.to(torch.int32)
    return torch.where(dygs, torch.cumsum(d, 0, dtype=torch.int32) - 1, torch.full_like(d, -1))


def rasterize_gaussians_raw(xyz, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, theta, rho,
                            raster_settings, mask=None, dx=None, ds=None, dr=None, dyn_slot=None):
    e = _empty()
    return _RasterizeGaussiansRaw.apply(xyz, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw,
                                        theta, rho, e if dx is None else dx, e if ds is None else ds, e if dr is None else dr,
                                        raster_settings, mask, dyn_slot)


class FusedGaussianRasterizer(nn.Module):
    """`GaussianRasterizer` fed with the model's raw parameters.  In `render()` it replaces

        rasterizer(means3D=pc.get_xyz, means2D=..., shs=pc.get_features, opacities=pc.get_opacity,
                   scales=pc.get_scaling, rotations=pc.get_rotation, theta=..., rho=...)
    by
        FusedGaussianRasterizer(raster_settings)(xyz=pc._xyz, means2D=..., features_dc=pc._features_dc,
                   features_rest=pc._features_rest, opacity_raw=pc._opacity, scaling_raw=pc._scaling,
                   rotation_raw=pc._rotation, theta=..., rho=...)

    and returns the same 5-tuple; gradients arrive directly on the raw parameters."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, xyz, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, theta=None, rho=None,
                mask=None, dx=None, ds=None, dr=None, dyn_slot=None):
        """`mask` (bool / uint8 [P]): Gaussians with mask == 0 are skipped inside the kernels -- the in-place version of
        `render(mask=...)`; `radii` / `n_touched` keep the FULL length P (the reference returns the gathered length).
        `dx, ds, dr` ([Pd,3], [Pd,3], [Pd,4]) + `dyn_slot` (`dynamic_slots(pc.dygs)`): the per-dynamic-Gaussian offsets of
        `render(dx=, ds=, dr=)`, added after the activations like the reference does."""
        if features_rest is None:
            features_rest = _empty()
        if theta is None:
            theta = _empty()
        if rho is None:
            rho = _empty()
        if (dx is not None or ds is not None or dr is not None) and dyn_slot is None:
            raise Exception("dx / ds / dr need dyn_slot = dynamic_slots(dygs)")
        return rasterize_gaussians_raw(xyz, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw,
                                       theta, rho, self.raster_settings, mask=mask, dx=dx, ds=ds, dr=dr, dyn_slot=dyn_slot)


# ----------------------------------------------------------------------------------------------
# inspection hook used by the parity tests (not part of the reference surface)
# ----------------------------------------------------------------------------------------------
def rasterize_gaussians_with_state(raster_settings, means3D, opacities, shs=None, colors_precomp=None, scales=None,
                                   rotations=None, cov3D_precomp=None):
    """Forward only (no autograd); returns the 5 outputs plus the saved integer state:
    ``num_rendered``, ``point_list`` (N,), ``ranges`` (tiles,2), ``n_contrib`` (H,W), ``final_T`` (H,W)."""
    e = _empty()
    with torch.no_grad():
        color, radii, depth, opacity, n_touched, st = _forward_impl(
            means3D, shs if shs is not None else e, colors_precomp if colors_precomp is not None else e, opacities,
            scales if scales is not None else e, rotations if rotations is not None else e,
            cov3D_precomp if cov3D_precomp is not None else e, raster_settings)
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    P, N = st["P"], st["N"]
    info = dict(num_rendered=N)
    if P:
        lay = _layout(P, W, H, st["capacity"])
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        img, binning, geom = st["img"], st["binning"], st["geom"]
        info["point_list"] = binning[lay.bin_point_list: lay.bin_point_list + 4 * N].view(torch.int32).clone()
        info["ranges"] = img[lay.img_ranges: lay.img_ranges + 8 * tiles].view(torch.int32).view(tiles, 2).clone()
        info["n_contrib"] = img[lay.img_n_contrib: lay.img_n_contrib + 4 * H * W].view(torch.int32).view(H, W).clone()
        info["final_T"] = img[lay.img_final_T: lay.img_final_T + 4 * H * W].view(torch.float32).view(H, W).clone()
        info["rec"] = geom[lay.geom_rec: lay.geom_rec + 48 * P].view(torch.float32).view(P, 12).clone()
    return (color, radii, depth, opacity, n_touched), info
