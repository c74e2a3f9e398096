"""The C-ABI library loads and exports every symbol include/g4r.h declares; sizing + argument validation work
without a GPU (no kernel is launched here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "g4r.h")
LIB = os.path.join(ROOT, "4dgs-slam_b200", "diff_gaussian_rasterization", "libg4r.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g4r_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("g4r_forward_project", "g4r_forward_render", "g4r_wait_num_rendered", "g4r_backward", "g4r_mark_visible",
              "g4r_geom_bytes", "g4r_image_bytes", "g4r_binning_bytes", "g4r_backward_scratch_bytes", "g4r_context_create",
              "g4r_context_destroy", "g4r_last_error", "g4r_version", "g4r_layout", "g4r_sort_scratch_bytes", "g4r_overflow_status"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_sizes():
    lib = ctypes.CDLL(LIB)
    assert lib.g4r_version() == 5
    for f in (lib.g4r_geom_bytes, lib.g4r_binning_bytes, lib.g4r_sort_scratch_bytes, lib.g4r_backward_scratch_bytes, lib.g4r_image_bytes):
        f.restype = ctypes.c_size_t
    lib.g4r_sort_scratch_bytes.argtypes = [ctypes.c_int64]
    lib.g4r_geom_bytes.argtypes = [ctypes.c_int32]
    lib.g4r_binning_bytes.argtypes = [ctypes.c_int64]
    lib.g4r_backward_scratch_bytes.argtypes = [ctypes.c_int32]
    lib.g4r_image_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    # 48-byte splat record + 1 clamp byte per Gaussian; per instance 4 bytes saved for backward (sorted ids) + 16 bytes of
    # forward-only sort scratch; 48 bytes of accumulators
    assert lib.g4r_geom_bytes(1000) >= 49 * 1000
    assert lib.g4r_geom_bytes(500_000) < 52 * 500_000
    assert 4 * 1_000_000 <= lib.g4r_binning_bytes(1_000_000) < 5 * 1_000_000       # only the sorted id list is saved
    assert lib.g4r_sort_scratch_bytes(1_000_000) >= 16 * 1_000_000
    assert lib.g4r_backward_scratch_bytes(1000) >= 48 * 1000
    assert lib.g4r_image_bytes(640, 480) >= 8 * 640 * 480 + 16 * 1200
    assert lib.g4r_geom_bytes(0) > 0 and lib.g4r_binning_bytes(0) > 0
    prev = 0
    for cap in (0, 1, 10, 1000, 10**6, 10**8):
        b = lib.g4r_binning_bytes(cap)
        assert b >= prev
        prev = b


def test_bad_arguments_are_reported_not_crashed():
    lib = ctypes.CDLL(LIB)
    lib.g4r_last_error.restype = ctypes.c_char_p
    rc = lib.g4r_forward_project(None, None, None, None, None, None, None, None)
    assert rc == -1 and b"frame" in lib.g4r_last_error()
    rc = lib.g4r_backward(None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"frame" in lib.g4r_last_error()
    lib.g4r_mark_visible.argtypes = [ctypes.c_int32] + [ctypes.c_void_p] * 5
    assert lib.g4r_mark_visible(-5, None, None, None, None, None) == -1
    assert lib.g4r_mark_visible(0, None, None, None, None, None) == 0
    lib.g4r_wait_num_rendered.restype = ctypes.c_int64
    assert lib.g4r_wait_num_rendered(None) == -1


def test_python_struct_mirrors_match_the_c_layout():
    import diff_gaussian_rasterization as dgr
    sizes = (ctypes.c_int32 * 5)()
    dgr._lib.g4r_struct_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(c) for c in (dgr._Frame, dgr._Gaussians, dgr._ForwardOut, dgr._BackwardIO, dgr._Layout)]
    assert [n for n, _ in dgr._Frame._fields_][-4:] == ["tile_rank", "tile_world", "tile_row_begin", "tile_row_end"]


def test_layout_offsets_are_aligned_and_ordered():
    from diff_gaussian_rasterization import _Layout, _lib
    lay = _Layout()
    assert _lib.g4r_layout(1000, 640, 480, 5000, ctypes.byref(lay)) == 0
    for name, _ in _Layout._fields_:
        assert getattr(lay, name) % 16 == 0, name
    assert lay.bin_point_list == 0 and lay.bin_pairs == 0          # offsets inside `binning` and inside `sort_scratch`
    assert lay.img_n_contrib - lay.img_final_T >= 4 * 640 * 480


def test_cpp_host_extension_matches_the_library():
    """csrc/host/g4r_torch.cpp (built by __graft_entry__.build_host, optional): compiled against the same g4r.h as libg4r.so --
    ABI version and struct sizes agree -- and it refuses CPU tensors with the same message as the Python host side."""
    import ctypes

    import torch

    import diff_gaussian_rasterization as dgr
    assert dgr.host_backend() in ("cpp", "python")
    if dgr.host_backend() != "cpp":
        pytest.skip("_g4r_host.so not built")
    host = dgr._host
    sizes = (ctypes.c_int32 * 5)()
    dgr._lib.g4r_struct_sizes(sizes)
    assert host.abi_version() == dgr._lib.g4r_version() == 5
    assert list(host.struct_sizes()) == list(sizes)[:4]
    e = torch.empty(0)
    args = (e, e, torch.zeros(4, 1), e, e, e, e, e, torch.ones(3), torch.eye(4), torch.eye(4), torch.eye(4), torch.zeros(3), 48, 64, 0.5, 0.5, 1.0, 0, False, 0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        host.rasterize(torch.zeros(4, 3), torch.zeros(4, 3), *args)
    with pytest.raises(RuntimeError, match=r"must have dimensions \(num_points, 3\)"):
        host.rasterize(torch.zeros(4, 2), torch.zeros(4, 3), *args)
