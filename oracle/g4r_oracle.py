"""ctypes/numpy front-end of the CPU oracle (oracle/g4r_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this module; the
product package (``4dgs-slam_b200/diff_gaussian_rasterization``) never does.

``Oracle("f32")`` follows the reference's float32 arithmetic (bit-faithful on radii / tile rectangles / depth
keys); ``Oracle("f64")`` is the same algorithm in double precision, used for finite-difference checks.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the oracle libraries with the committed Makefile (gcc, -ffp-contract=off)."""
    libs = ("libg4r_oracle_f32.so", "libg4r_oracle_f64.so", "libknn_oracle.so", "libknn_search_host.so")
    srcs = [os.path.join(_HERE, "g4r_oracle.c"), os.path.join(_HERE, "knn_oracle.c"),
            os.path.join(_HERE, "..", "tests", "native", "knn_search_host.cpp"),
            os.path.join(_HERE, "..", "4dgs-slam_b200", "csrc", "knn_search.cuh")]
    need = force or not all(os.path.exists(os.path.join(_BUILD, f)) for f in libs)
    if not need:
        newest = max(os.path.getmtime(s) for s in srcs)
        need = any(os.path.getmtime(os.path.join(_BUILD, f)) < newest for f in libs)
    if need:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)


def knn_mean_dist2(points: np.ndarray) -> np.ndarray:
    """distCUDA2 restated as a brute force (oracle/knn_oracle.c): float32 [P] from float32 [P,3]."""
    build()
    lib = ctypes.CDLL(os.path.join(_BUILD, "libknn_oracle.so"))
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    out = np.empty(pts.shape[0], np.float32)
    lib.knn_oracle_mean_dist2(ctypes.c_int32(pts.shape[0]), ctypes.c_void_p(pts.ctypes.data), ctypes.c_void_p(out.ctypes.data))
    return out


def _scene_struct(real):
    class OracleScene(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int32) for n in ("P", "D", "M", "W", "H")] + \
                   [(n, real) for n in ("tan_fovx", "tan_fovy", "scale_modifier")] + \
                   [(n, ctypes.c_void_p) for n in ("bg", "viewmatrix", "projmatrix", "projmatrix_raw", "campos", "means3D",
                                                   "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp")]
    return OracleScene


class _Geom(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped", "radii", "tiles_touched")]


class _Grads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dcolors", "dL_dcov3D", "dL_dshs",
                                               "dL_dscales", "dL_drots", "dL_dtau_rows", "dL_dtau")]


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Oracle:
    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        build()
        self.precision = precision
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.real = ctypes.c_float if precision == "f32" else ctypes.c_double
        self.lib = ctypes.CDLL(os.path.join(_BUILD, f"libg4r_oracle_{precision}.so"))
        self.Scene = _scene_struct(self.real)
        s = "_" + precision
        self._project = getattr(self.lib, "oracle_project" + s)
        self._bin = getattr(self.lib, "oracle_bin" + s)
        self._bin.restype = ctypes.c_int64
        self._free = getattr(self.lib, "oracle_free" + s)
        self._free.argtypes = [ctypes.c_void_p]
        self._composite = getattr(self.lib, "oracle_composite" + s)
        self._composite_bw = getattr(self.lib, "oracle_composite_bw" + s)
        self._gaussian_bw = getattr(self.lib, "oracle_gaussian_bw" + s)
        self._mark_visible = getattr(self.lib, "oracle_mark_visible" + s)

    # ------------------------------------------------------------------------------------------
    def _arr(self, a, shape=None):
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, dtype=self.dtype))
        if shape is not None:
            a = a.reshape(shape)
        return a

    def _pack(self, sc: dict):
        """sc: dict with the keys of tools.scenes.Scene (numpy or torch CPU tensors)."""
        def get(k):
            v = sc.get(k)
            if v is None:
                return None
            if hasattr(v, "detach"):
                v = v.detach().cpu().numpy()
            return self._arr(v)

        keep = {k: get(k) for k in ("bg", "viewmatrix", "projmatrix", "projmatrix_raw", "campos", "means3D", "opacities", "shs",
                                    "colors_precomp", "scales", "rotations", "cov3D_precomp")}
        P = int(keep["means3D"].shape[0])
        M = int(keep["shs"].shape[1]) if keep["shs"] is not None else 0
        S = self.Scene()
        S.P, S.D, S.M, S.W, S.H = P, int(sc["sh_degree"]), M, int(sc["W"]), int(sc["H"])
        # the reference receives tanfov / scale_modifier as C floats (pybind double -> float)
        S.tan_fovx = float(np.float32(sc["tanfovx"])) if self.precision == "f32" else float(sc["tanfovx"])
        S.tan_fovy = float(np.float32(sc["tanfovy"])) if self.precision == "f32" else float(sc["tanfovy"])
        S.scale_modifier = float(sc.get("scale_modifier", 1.0))
        for k, v in keep.items():
            setattr(S, k, _p(v))
        return S, keep, P, M

    # ------------------------------------------------------------------------------------------
    def forward(self, sc: dict) -> dict:
        """Full forward: projection, binning, composite.  Returns outputs + every intermediate."""
        S, keep, P, M = self._pack(sc)
        W, H = S.W, S.H
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        dt = self.dtype
        geom = dict(depths=np.zeros(P, dt), means2D=np.zeros((P, 2), dt), cov3D=np.zeros((P, 6), dt),
                    conic_opacity=np.zeros((P, 4), dt), rgb=np.zeros((P, 3), dt), clamped=np.zeros((P, 3), np.uint8),
                    radii=np.zeros(P, np.int32), tiles_touched=np.zeros(P, np.uint32))
        G = _Geom(*[_p(geom[k]) for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped", "radii", "tiles_touched")])
        self._project(ctypes.byref(S), ctypes.byref(G))
        pl_ptr = ctypes.c_void_p()
        ranges = np.zeros((tiles, 2), np.uint32)
        N = int(self._bin(ctypes.byref(S), ctypes.byref(G), ctypes.byref(pl_ptr), ctypes.c_void_p(_p(ranges))))
        point_list = np.ctypeslib.as_array(ctypes.cast(pl_ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(max(N, 1),))[:N].copy()
        self._free(pl_ptr)
        color = np.zeros((3, H, W), dt)
        depth = np.zeros((1, H, W), dt)
        opacity = np.zeros((1, H, W), dt)
        final_T = np.zeros((H, W), dt)
        n_contrib = np.zeros((H, W), np.uint32)
        n_touched = np.zeros(P, np.int32)
        pl = np.ascontiguousarray(point_list if N else np.zeros(1, np.uint32))
        self._composite(ctypes.byref(S), ctypes.byref(G), ctypes.c_void_p(_p(pl)), ctypes.c_void_p(_p(ranges)),
                        ctypes.c_void_p(_p(color)), ctypes.c_void_p(_p(depth)), ctypes.c_void_p(_p(opacity)),
                        ctypes.c_void_p(_p(final_T)), ctypes.c_void_p(_p(n_contrib)), ctypes.c_void_p(_p(n_touched)))
        out = dict(color=color, depth=depth, opacity=opacity, radii=geom["radii"], n_touched=n_touched, num_rendered=N,
                   point_list=point_list, ranges=ranges, final_T=final_T, n_contrib=n_contrib, **{k: geom[k] for k in geom if k != "radii"})
        out["_state"] = (S, keep, G, geom, pl, ranges, P, M)
        return out

    def backward(self, fwd: dict, grad_color, grad_depth) -> dict:
        """Analytic backward of the reference, given a forward() result and upstream image gradients."""
        S, keep, G, geom, pl, ranges, P, M = fwd["_state"]
        dt = self.dtype
        gc = self._arr(grad_color.detach().cpu().numpy() if hasattr(grad_color, "detach") else grad_color)
        gd = self._arr(grad_depth.detach().cpu().numpy() if hasattr(grad_depth, "detach") else grad_depth)
        acc = np.zeros((P, 10), np.float64)
        self._composite_bw(ctypes.byref(S), ctypes.byref(G), ctypes.c_void_p(_p(pl)), ctypes.c_void_p(_p(ranges)),
                           ctypes.c_void_p(_p(fwd["final_T"])), ctypes.c_void_p(_p(fwd["n_contrib"])), ctypes.c_void_p(_p(gc)),
                           ctypes.c_void_p(_p(gd)), ctypes.c_void_p(_p(acc)))
        has_sh, has_scale = keep["shs"] is not None, keep["scales"] is not None
        g = dict(dL_dmeans3D=np.zeros((P, 3), dt), dL_dmeans2D=np.zeros((P, 3), dt), dL_dopacity=np.zeros(P, dt),
                 dL_dcolors=np.zeros((P, 3), dt), dL_dcov3D=np.zeros((P, 6), dt),
                 dL_dshs=np.zeros((P, M, 3), dt) if has_sh else None,
                 dL_dscales=np.zeros((P, 3), dt) if has_scale else None, dL_drots=np.zeros((P, 4), dt) if has_scale else None,
                 dL_dtau_rows=np.zeros((P, 6), dt), dL_dtau=np.zeros(6, dt))
        GR = _Grads(*[_p(g[k]) for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dcolors", "dL_dcov3D", "dL_dshs", "dL_dscales",
                                         "dL_drots", "dL_dtau_rows", "dL_dtau")])
        self._gaussian_bw(ctypes.byref(S), ctypes.byref(G), ctypes.c_void_p(_p(acc)), ctypes.byref(GR))
        g["acc"] = acc
        g["grad_rho"] = g["dL_dtau"][:3].copy()
        g["grad_theta"] = g["dL_dtau"][3:].copy()
        return g

    def mark_visible(self, means3D, viewmatrix) -> np.ndarray:
        m = self._arr(means3D)
        v = self._arr(viewmatrix)
        out = np.zeros(m.shape[0], np.uint8)
        self._mark_visible(ctypes.c_int32(m.shape[0]), ctypes.c_void_p(_p(m)), ctypes.c_void_p(_p(v)), ctypes.c_void_p(_p(out)))
        return out.astype(bool)


def scene_dict(scene) -> dict:
    """tools.scenes.Scene -> plain dict understood by Oracle.forward."""
    return dict(scene.__dict__)


# ---------------------------------------------------------------------------------------------------------------------
# Raw-parameter mode (include/g4r.h G4R_ACT_RAW).  Restates the activation prelude of the
# reference's render() (gaussian_splatting/gaussian_renderer/__init__.py:108-131 with
# gaussian_splatting/scene/gaussian_model.py:100-128: get_scaling = exp, get_opacity = sigmoid, get_rotation = normalize,
# get_features = cat(dc, rest)) and the chain rule autograd applies to it, in numpy.  Pinned by
# tests/test_raw_activation.py against torch's own autograd through exactly those torch calls.
# ---------------------------------------------------------------------------------------------------------------------
def activate_raw(opacity_raw, features_dc, features_rest, scaling_raw, rotation_raw, dtype=np.float32) -> dict:
    """Raw GaussianModel parameters -> the activated tensors the reference rasterizer receives."""
    f = lambda a: None if a is None else np.asarray(a, dtype=dtype)
    o, dc, rest, s, q = f(opacity_raw), f(features_dc), f(features_rest), f(scaling_raw), f(rotation_raw)
    one = dtype(1.0)
    scales = np.exp(s)
    if scales.shape[1] == 1:
        scales = np.repeat(scales, 3, axis=1)
    # |q|^2 accumulated r, x, y, z with one rounding per step, like g4r_quat_norm (fma chain) up to the fma's single rounding
    n = np.maximum(np.sqrt((q.astype(np.float64) ** 2).sum(1)).astype(dtype), dtype(1e-12))[:, None]
    return dict(opacities=(one / (one + np.exp(-o))).astype(dtype), scales=scales.astype(dtype), rotations=(q / n).astype(dtype),
                shs=dc if rest is None or rest.shape[1] == 0 else np.concatenate([dc, rest], axis=1), rot_norm=n)


def raw_chain_rule(act: dict, grads: dict, scale_dim: int = 3) -> dict:
    """Gradients w.r.t. the activated tensors (Oracle.backward output) -> gradients w.r.t. the raw parameters."""
    o = act["opacities"].reshape(-1).astype(np.float64)
    s = act["scales"].astype(np.float64)
    q = act["rotations"].astype(np.float64)
    n = act["rot_norm"].astype(np.float64)
    gs = grads["dL_dscales"].astype(np.float64) * s
    if scale_dim == 1:
        gs = gs.sum(1, keepdims=True)
    gq = grads["dL_drots"].astype(np.float64)
    gq = (gq - q * (q * gq).sum(1, keepdims=True)) / n
    gsh = grads["dL_dshs"]
    return dict(dL_dopacity_raw=grads["dL_dopacity"].astype(np.float64).reshape(-1) * o * (1.0 - o), dL_dscaling_raw=gs,
                dL_drotation_raw=gq, dL_dfeatures_dc=gsh[:, :1], dL_dfeatures_rest=gsh[:, 1:])


# ---------------------------------------------------------------------------------------------------------------------
# fused SLAM losses: numpy restatement of utils/slam_utils.py, pinned to tests/golden/slam_loss.npz
# (vectors produced by the reference's own functions + torch autograd, tools/make_loss_golden.py)
# ---------------------------------------------------------------------------------------------------------------------
def slam_loss_ref(mode: str, image, depth, gt_image, gt_depth, opacity=None, exposure=(0.0, 0.0), motion_mask=None, grad_mask=None,
                  alpha: float = 0.95, thr: float = 0.01) -> dict:
    """get_loss_tracking -> get_loss_tracking_rgbd (slam_utils.py:57-173; mode "tracking") or get_loss_mapping ->
    get_loss_mapping_rgbd (:252-364, static non-split branch; mode "mapping") in float64, with the closed-form gradients
    w.r.t. the rendered colour, depth and the exposure pair.  `motion_mask=None` = the reference's "mask not applied" case."""
    f = np.float64
    image, depth = np.asarray(image, f).reshape(3, -1), np.asarray(depth, f).reshape(-1)
    gt, gd = np.asarray(gt_image, np.float32).reshape(3, -1), np.asarray(gt_depth, f).reshape(-1)
    X = depth.size
    a, b = float(exposure[0]), float(exposure[1])
    ea = np.exp(a)
    mm = np.ones(X, f) if motion_mask is None else (np.asarray(motion_mask).reshape(-1) != 0).astype(f)
    m = ((gt[0] + gt[1]) + gt[2] > np.float32(thr)).astype(f) * mm                  # gt_image.sum(dim=0) > rgb_boundary_threshold (float32)
    gt = gt.astype(f)
    if mode == "tracking":
        o = np.asarray(opacity, f).reshape(-1)
        if grad_mask is not None:
            m = m * (np.asarray(grad_mask).reshape(-1) != 0)
        w = o
        md = ((gd > 0.01) & (gd < 1000.0) & (np.asarray(opacity, np.float32).reshape(-1) > np.float32(0.95))).astype(f) * mm
    else:
        w = np.ones(X, f)
        md = ((gd > 0.01) & (gd < 10000.0)).astype(f) * mm
    diff = (ea * image + b) * m - gt * m
    s = np.sign(diff) * m * w
    dd = depth * md - gd * md
    k_rgb, k_d = alpha / (3.0 * X), (1.0 - alpha) / X
    return dict(loss=k_rgb * (w * np.abs(diff)).sum() + k_d * np.abs(dd).sum(), d_image=(k_rgb * s * ea).reshape(np.shape(image)),
                d_depth=k_d * np.sign(dd) * md, d_exposure=np.array([k_rgb * (s * ea * image).sum(), k_rgb * s.sum()]))


def prelude_ref(xyz, scaling_act, rotation_act, dygs, dx=None, ds=None, dr=None, mask=None) -> dict:
    """The remaining prelude of render() (gaussian_renderer/__init__.py:159-191) in numpy: dynamic offsets scattered into the rows
    of the dynamic Gaussians and added to positions / ACTIVATED scales / ACTIVATED rotations, then the static-mask gather.
    Returns the tensors the rasterizer receives plus the row indices kept by the mask."""
    xyz, sc, rot = np.array(xyz, copy=True), np.array(scaling_act, copy=True), np.array(rotation_act, copy=True)
    dy = np.asarray(dygs).astype(bool)
    if dx is not None and ds is not None and dr is not None:
        xyz[dy] += np.asarray(dx)
        sc[dy] += np.asarray(ds)
        rot[dy] += np.asarray(dr)
    keep = np.arange(xyz.shape[0]) if mask is None else np.nonzero(np.asarray(mask).astype(bool))[0]
    return dict(means3D=xyz[keep], scales=sc[keep], rotations=rot[keep], keep=keep)


# ---------------------------------------------------------------------------------------------------------------------
# control-node warp of the deformation step: torch restatement (CPU, any dtype, autograd gives the
# gradients) of ControlNodeWarp.forward, utils/time_utils.py:1192-1275, with cal_nn_weight :981-1015 and quaternion_to_matrix
# :115-132.  PARITY UNPINNED against the reference itself: the reference calls pytorch3d.ops.knn_points (:998), an un-vendored,
# un-pinned dependency (requirements.txt:19, "git+https://github.com/facebookresearch/pytorch3d.git") that is absent here, so the
# class cannot be imported.  knn_points' published contract -- squared Euclidean distances to the K nearest points, ascending,
# with their indices -- is restated as an explicit distance matrix + stable sort; everything after it follows the reference's
# statements line by line (cited).  tests/test_deform.py additionally checks the closed-form gradients csrc/warp.cu uses against
# autograd of this restatement in float64.
# ---------------------------------------------------------------------------------------------------------------------
def control_node_warp_ref(x, nodes, log_radius, weight_logit, node_attrs, motion_mask=None, K=3, d_rot_as_res=True, local_frame=True):
    import torch
    x = x.detach()                                                        # :1196
    n3 = nodes[..., :3].detach()                                          # :994
    diff = x[:, None, :] - n3[None, :, :]
    dist = ((diff * diff)[..., 0] + (diff * diff)[..., 1]) + (diff * diff)[..., 2]
    nn_dist, nn_idx = torch.sort(dist, dim=1, stable=True)                # knn_points: ascending squared distances
    nn_dist, nn_idx = nn_dist[:, :K], nn_idx[:, :K]
    nn_radius = torch.exp(log_radius).reshape(-1)[nn_idx]                 # :893, :1001
    nn_weight = torch.exp(-nn_dist / (2 * nn_radius ** 2))                # :1002
    if weight_logit is not None:
        nn_weight = nn_weight * torch.sigmoid(weight_logit).reshape(-1, 1)[nn_idx][..., 0]   # :897, :1004-1005
    nn_weight = nn_weight + 1e-7                                          # :1006
    nn_weight = nn_weight / nn_weight.sum(dim=-1, keepdim=True)           # :1007
    mask = 1.0 if motion_mask is None else motion_mask.reshape(-1, 1)
    rot_bias = torch.tensor([1.0, 0, 0, 0], dtype=x.dtype)
    node_trans, node_rot, node_scale = node_attrs["d_xyz"], node_attrs["d_rotation"], node_attrs["d_scaling"]
    if local_frame:                                                       # :1208-1214
        q = node_attrs["local_rotation"] + rot_bias
        r, i, j, k = torch.unbind(q, -1)
        two_s = 2.0 / (q * q).sum(-1)
        R = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r), two_s * (i * j + k * r),
                         1 - two_s * (i * i + k * k), two_s * (j * k - i * r), two_s * (i * k - j * r), two_s * (j * k + i * r),
                         1 - two_s * (i * i + j * j)), -1).reshape(-1, 3, 3)
        nn_nodes = n3[nn_idx]
        Ax = torch.einsum("nkab,nkb->nka", R[nn_idx], x[:, None] - nn_nodes) + nn_nodes + node_trans[nn_idx]
        translate = (Ax * nn_weight[..., None]).sum(dim=1) - x
    else:
        translate = (node_trans[nn_idx] * nn_weight[..., None]).sum(dim=1)            # :1216
    translate = translate * mask                                                        # :1217
    if not d_rot_as_res:
        rotation = (((node_rot + rot_bias)[nn_idx] * nn_weight[..., None]).sum(dim=1) - rot_bias) * mask + rot_bias   # :1222, :1232
    else:
        rotation = (node_rot[nn_idx] * nn_weight[..., None]).sum(dim=1) * mask          # :1251-1252
    scale = (node_scale[nn_idx] * nn_weight[..., None]).sum(dim=1) * mask               # :1247 / :1253
    return dict(d_xyz=translate, d_rotation=rotation, d_scaling=scale, nn_weight=nn_weight, nn_dist=nn_dist, nn_idx=nn_idx)
