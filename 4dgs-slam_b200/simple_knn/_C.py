"""``simple_knn._C`` of the reference exposes one function, ``distCUDA2`` (submodules/simple-knn/ext.cpp, spatial.cu:15-26)."""
import ctypes

import torch

from diff_gaussian_rasterization import _check, _lib

__all__ = ["distCUDA2"]

_lib.g4r_knn_scratch_bytes.restype = ctypes.c_size_t
_lib.g4r_knn_scratch_bytes.argtypes = [ctypes.c_int32]
_lib.g4r_knn_mean_dist2.restype = ctypes.c_int
_lib.g4r_knn_mean_dist2.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]

_bytes = {}


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance of every point to its 3 nearest neighbours; ``points`` is a CUDA float tensor [P, 3], the result a
    float32 tensor [P] on the same device (spatial.cu:15-26).  Enqueued on the current stream; no host synchronisation."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: points must live on a CUDA device; there is no CPU path")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError(f"distCUDA2: points must have shape [P, 3], got {tuple(points.shape)}")
    P = points.shape[0]
    pts = points.detach()
    if pts.dtype is not torch.float32 or not pts.is_contiguous():
        pts = pts.to(torch.float32).contiguous()
    with torch.cuda.device(points.device):
        means = torch.full((P,), 0.0, dtype=torch.float32, device=points.device)        # spatial.cu:21
        if P == 0:
            return means
        nbytes = _bytes.get(P)
        if nbytes is None:
            nbytes = _bytes[P] = int(_lib.g4r_knn_scratch_bytes(P))
            if nbytes == 0:
                _check(-1)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
        _check(_lib.g4r_knn_mean_dist2(P, pts.data_ptr(), means.data_ptr(), scratch.data_ptr(), nbytes,
                                       torch.cuda.current_stream(points.device).cuda_stream))
    return means
