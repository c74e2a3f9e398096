"""Synthetic Gaussian clouds + cameras for parity tests, goldens and bench.py.

Everything is generated on the CPU with a seeded ``torch.Generator`` so every machine sees the same
float32 bits.  Camera matrices are built the way the reference builds them:
``getWorld2View2`` / ``getProjectionMatrix2`` (gaussian_splatting/utils/graphics_utils.py:33-46,72-93) and the
``Camera`` properties (utils/camera_utils.py:124-148): viewmatrix = W2C^T, projmatrix = W2C^T @ P^T,
projmatrix_raw = P^T, campos = inv(viewmatrix)[3,:3].
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch


@dataclass
class Scene:
    """One frame's worth of rasterizer inputs (CPU float32 tensors)."""
    name: str
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    campos: torch.Tensor
    sh_degree: int
    scale_modifier: float
    means3D: torch.Tensor
    opacities: torch.Tensor                      # (P,1)
    shs: Optional[torch.Tensor] = None           # (P,M,3)
    colors_precomp: Optional[torch.Tensor] = None
    scales: Optional[torch.Tensor] = None
    rotations: Optional[torch.Tensor] = None
    cov3D_precomp: Optional[torch.Tensor] = None
    grad_color: Optional[torch.Tensor] = None    # upstream dL/dcolor (3,H,W)
    grad_depth: Optional[torch.Tensor] = None    # upstream dL/ddepth (1,H,W)
    meta: dict = field(default_factory=dict)

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])

    def to(self, device) -> "Scene":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Scene(**kw)


def broadcast_scene(sc: Scene, src: int = 0, group=None) -> Scene:
    """Multi-rank runs: every rank takes rank `src`'s tensors, so that all ranks hold the bit-identical scene (the CPU
    generator's float64 transcendental / reduction kernels are not guaranteed to round identically in every process)."""
    import torch.distributed as dist
    for k, v in sc.__dict__.items():
        if isinstance(v, torch.Tensor):
            dist.broadcast(v, src=src, group=group)
    return sc


def _se3_exp(rho: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """SE(3) exponential (float64) used only to produce a generic, non-axis-aligned test pose."""
    W = torch.tensor([[0, -theta[2], theta[1]], [theta[2], 0, -theta[0]], [-theta[1], theta[0], 0]], dtype=torch.float64)
    a = float(torch.linalg.norm(theta))
    I = torch.eye(3, dtype=torch.float64)
    if a < 1e-8:
        R, Vm = I + W, I + 0.5 * W
    else:
        R = I + math.sin(a) / a * W + (1 - math.cos(a)) / a**2 * (W @ W)
        Vm = I + (1 - math.cos(a)) / a**2 * W + (a - math.sin(a)) / a**3 * (W @ W)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = R
    T[:3, 3] = Vm @ rho.double()
    return T


def camera_matrices(W, H, fx, fy, cx, cy, R_w2c: torch.Tensor, t_w2c: torch.Tensor, znear=0.01, zfar=100.0):
    """Reference conventions, float32 like the reference (graphics_utils.py:33-46,72-93; camera_utils.py:124-148)."""
    Rt = torch.zeros((4, 4), dtype=torch.float32)
    Rt[:3, :3] = R_w2c.float()
    Rt[:3, 3] = t_w2c.float()
    Rt[3, 3] = 1.0
    Rt = torch.linalg.inv(torch.linalg.inv(Rt))          # getWorld2View2 round-trips through C2W
    viewmatrix = Rt.transpose(0, 1).contiguous()
    left = ((2 * cx - W) / W - 1.0) * W / 2.0
    right = ((2 * cx - W) / W + 1.0) * W / 2.0
    top = ((2 * cy - H) / H + 1.0) * H / 2.0
    bottom = ((2 * cy - H) / H - 1.0) * H / 2.0
    left, right = znear / fx * left, znear / fx * right
    top, bottom = znear / fy * top, znear / fy * bottom
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    projmatrix_raw = Pm.transpose(0, 1).contiguous()
    projmatrix = (viewmatrix.unsqueeze(0).bmm(projmatrix_raw.unsqueeze(0))).squeeze(0).contiguous()
    campos = viewmatrix.inverse()[3, :3].contiguous()
    FoVx = 2 * math.atan(W / (2 * fx))                    # utils/dataset.py:278-279
    FoVy = 2 * math.atan(H / (2 * fy))
    return dict(viewmatrix=viewmatrix, projmatrix=projmatrix, projmatrix_raw=projmatrix_raw, campos=campos,
                tanfovx=math.tan(FoVx * 0.5), tanfovy=math.tan(FoVy * 0.5))


def make_scene(P: int, W: int, H: int, sh_degree: int = 0, sh_coeffs: Optional[int] = None, seed: int = 0,
               off_centre: bool = True, posed: bool = True, colors_precomp: bool = False, cov3D_precomp: bool = False,
               scale_modifier: float = 1.0, px_min: float = 0.5, px_max: float = 3.0, name: str = "") -> Scene:
    """SURVEY.md section 8d cloud: depth U(0.5,6); pixel positions U(-5%,105%) of the image; anisotropic scales of
    LogUniform(px_min, px_max) pixels; random rotations; opacity sigmoid(N(0,1.5^2)); 1% of points behind the near plane."""
    gen = torch.Generator().manual_seed(seed)

    def U(n, a, b):
        return torch.rand(n, generator=gen, dtype=torch.float64) * (b - a) + a

    def Nrm(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float64)

    fx = fy = 0.8366 * W                                   # TUM: 535.4 / 640
    if off_centre:
        cx, cy = W / 2 + 0.1 * W / 640, H / 2 + 7.6 * H / 480
    else:
        cx, cy = W / 2, H / 2

    z = U(P, 0.5, 6.0)
    n_behind = max(1, P // 100) if P >= 8 else 0
    if n_behind:
        z[:n_behind] = U(n_behind, -1.0, 0.2)
    u = U(P, -0.05 * W, 1.05 * W)
    v = U(P, -0.05 * H, 1.05 * H)
    pc = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], dim=1)        # camera-space points

    if posed:
        T_w2c = _se3_exp(torch.tensor([0.31, -0.17, 0.23]), torch.tensor([0.12, -0.21, 0.07]))
    else:
        T_w2c = torch.eye(4, dtype=torch.float64)
    R, t = T_w2c[:3, :3], T_w2c[:3, 3]
    # world points R^T (pc - t), written element-wise: BLAS results depend on buffer alignment, which would make the cloud
    # differ in the last bit between processes (seen as 1-ulp image differences between ranks / runs)
    d = pc - t
    pw = torch.stack([d[:, 0] * R[0, j] + d[:, 1] * R[1, j] + d[:, 2] * R[2, j] for j in range(3)], dim=1)
    cam = camera_matrices(W, H, fx, fy, cx, cy, R, t)

    r_px = torch.exp(U(P * 3, math.log(px_min), math.log(px_max))).view(P, 3)
    scales = (z.abs().clamp_min(0.2) / fx).unsqueeze(1) * r_px
    q = Nrm(P, 4)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(Nrm(P, 1) * 1.5)
    M = sh_coeffs if sh_coeffs is not None else (sh_degree + 1) ** 2
    shs = torch.cat([Nrm(P, 1, 3), Nrm(P, max(M - 1, 0), 3) * 0.1], dim=1)[:, :M]
    grad_color = Nrm(3, H, W) / (W * H)
    grad_depth = Nrm(1, H, W) / (W * H)

    f32 = lambda a: a.float().contiguous()
    sc = Scene(name=name or f"P{P}_{W}x{H}_d{sh_degree}_s{seed}", W=W, H=H, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
               bg=torch.tensor([1.0, 1.0, 1.0]), viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"],
               projmatrix_raw=cam["projmatrix_raw"], campos=cam["campos"], sh_degree=sh_degree,
               scale_modifier=scale_modifier, means3D=f32(pw), opacities=f32(opacities),
               grad_color=f32(grad_color), grad_depth=f32(grad_depth),
               meta=dict(seed=seed, fx=fx, fy=fy, cx=cx, cy=cy, off_centre=off_centre, posed=posed))
    if colors_precomp:
        sc.colors_precomp = f32(torch.rand(P, 3, generator=gen, dtype=torch.float64))
    else:
        sc.shs = f32(shs)
    if cov3D_precomp:
        # symmetric PSD from the same scale/rotation, in float64 then rounded
        qr, qx, qy, qz = rotations.unbind(1)
        Rm = torch.stack([1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qr * qz), 2 * (qx * qz + qr * qy),
                          2 * (qx * qy + qr * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qr * qx),
                          2 * (qx * qz - qr * qy), 2 * (qy * qz + qr * qx), 1 - 2 * (qx * qx + qy * qy)], dim=1).view(P, 3, 3)
        L = Rm * (scales * scale_modifier).unsqueeze(1)
        S = L @ L.transpose(1, 2)
        sc.cov3D_precomp = f32(torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1))
    else:
        sc.scales = f32(scales)
        sc.rotations = f32(rotations)
    return sc


# ---------------------------------------------------------------------------------------------------------------------
# bit-reproducible scenes for the committed reference-output digests (tests/golden/digests_ref.json)
# ---------------------------------------------------------------------------------------------------------------------
def _splitmix_uniform(seed: int, stream: int, n: int) -> np.ndarray:
    """n doubles in [0, 1) from SplitMix64 on a counter: integer arithmetic only, so every machine produces the same bits."""
    with np.errstate(over="ignore"):
        x = (np.arange(1, n + 1, dtype=np.uint64) + np.uint64((seed * 1000003 + stream * 7919) & 0xFFFFFFFF) * np.uint64(0x632BE59BD9B4E019))
        x = x * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def make_scene_exact(P: int, W: int, H: int, sh_degree: int = 0, seed: int = 0, px_min: float = 0.5, px_max: float = 3.0,
                     name: str = "") -> Scene:
    """Same kind of cloud as make_scene, but built ONLY from operations that IEEE-754 rounds
    identically everywhere: integer hashing for the random numbers and element-wise + - * / sqrt in float64 (no exp / log /
    sin / BLAS / LAPACK, whose last bit depends on the library, the vector width and the buffer alignment).  The float32
    inputs -- and therefore the digests of the reference's outputs committed under tests/golden/ -- are reproducible on any
    machine.  Differences to make_scene: reciprocal-uniform instead of log-uniform splat sizes, Irwin-Hall(4) instead of
    Gaussian variates, uniform opacities (0.4 % below 1/255), a fixed rational camera rotation."""
    stream = [0]

    def U(n):
        stream[0] += 1
        return _splitmix_uniform(seed, stream[0], n)

    def Nrm(n):                                              # Irwin-Hall(4), unit variance
        return (U(n) + U(n) + U(n) + U(n) - 2.0) * 1.7320508075688772

    fx = fy = 0.8366 * W
    cx, cy = W / 2 + 0.1 * W / 640, H / 2 + 7.6 * H / 480
    z = 0.5 + 5.5 * U(P)
    n_behind = max(1, P // 100) if P >= 8 else 0
    if n_behind:
        z[:n_behind] = -1.0 + 1.2 * U(n_behind)
    u = (-0.05 + 1.1 * U(P)) * W
    v = (-0.05 + 1.1 * U(P)) * H
    pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], axis=1)
    # world -> camera: p_c = R p_w + t with a rational rotation (rows (1,-4,8)/9, (8,4,1)/9, (-4,7,4)/9) and t = (0.3, -0.2, 0.25)
    R = np.array([[1.0, -4.0, 8.0], [8.0, 4.0, 1.0], [-4.0, 7.0, 4.0]]) / 9.0
    t = np.array([0.3, -0.2, 0.25])
    d = pc - t
    pw = np.stack([d[:, 0] * R[0, j] + d[:, 1] * R[1, j] + d[:, 2] * R[2, j] for j in range(3)], axis=1)      # R^T (p_c - t)

    Rf, tf = R.astype(np.float32), t.astype(np.float32)
    view = np.zeros((4, 4), np.float32)                      # viewmatrix = [R t; 0 1]^T (reference layout, Appendix A.1)
    view[:3, :3] = Rf.T
    view[3, :3] = tf
    view[3, 3] = 1.0
    znear, zfar = 0.01, 100.0
    left = ((2 * cx - W) / W - 1.0) * W / 2.0
    right = ((2 * cx - W) / W + 1.0) * W / 2.0
    top = ((2 * cy - H) / H + 1.0) * H / 2.0
    bottom = ((2 * cy - H) / H - 1.0) * H / 2.0
    left, right, top, bottom = znear / fx * left, znear / fx * right, znear / fy * top, znear / fy * bottom
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    proj_raw = Pm.T.astype(np.float32)
    v64, p64 = view.astype(np.float64), proj_raw.astype(np.float64)
    proj = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            acc = 0.0
            for k in range(4):
                acc = acc + float(v64[i, k]) * float(p64[k, j])      # plain Python doubles: one rounding per operation
            proj[i, j] = acc
    campos = np.array([-(float(R[0, j]) * float(t[0]) + float(R[1, j]) * float(t[1]) + float(R[2, j]) * float(t[2])) for j in range(3)])

    r_px = (px_min * px_max) / (px_max - U(P * 3) * (px_max - px_min))
    scales = (np.maximum(np.abs(z), 0.2) / fx)[:, None] * r_px.reshape(P, 3)
    q = np.stack([Nrm(P) for _ in range(4)], axis=1)
    qn = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    qn = np.where(qn > 1e-12, qn, 1.0)
    rotations = q / qn[:, None]
    opacities = U(P).reshape(P, 1)
    M = (sh_degree + 1) ** 2
    shs = np.concatenate([Nrm(P * 3).reshape(P, 1, 3), 0.1 * Nrm(P * max(M - 1, 0) * 3).reshape(P, max(M - 1, 0), 3)], axis=1)
    grad_color = Nrm(3 * H * W).reshape(3, H, W) / (W * H)
    grad_depth = Nrm(H * W).reshape(1, H, W) / (W * H)

    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))
    return Scene(name=name or f"X{P}_{W}x{H}_d{sh_degree}_s{seed}", W=W, H=H, tanfovx=W / (2.0 * fx), tanfovy=H / (2.0 * fy),
                 bg=torch.tensor([1.0, 1.0, 1.0]), viewmatrix=f32(view), projmatrix=f32(proj), projmatrix_raw=f32(proj_raw),
                 campos=f32(campos), sh_degree=sh_degree, scale_modifier=1.0, means3D=f32(pw), opacities=f32(opacities),
                 shs=f32(shs), scales=f32(scales), rotations=f32(rotations), grad_color=f32(grad_color), grad_depth=f32(grad_depth),
                 meta=dict(seed=seed, exact=True))


EXACT_CONFIGS = {        # BASELINE.json config sizes on the bit-reproducible generator
    "X1": dict(P=10_000, W=320, H=240, sh_degree=0),
    "X2": dict(P=100_000, W=640, H=480, sh_degree=3),
    "X3": dict(P=500_000, W=640, H=480, sh_degree=0),
    "X4": dict(P=2_000_000, W=1280, H=960, sh_degree=0),
}


LARGE_CONFIGS = {        # beyond BASELINE.json: a frame large enough for the Gaussian-sharded render to pay (no reference digests)
    "X5": dict(P=8_000_000, W=1920, H=1088, sh_degree=0),
}


def exact_scene(name: str, seed: int = 0) -> Scene:
    return make_scene_exact(seed=seed, name=name, **(EXACT_CONFIGS.get(name) or LARGE_CONFIGS[name]))


def input_digest(sc: Scene) -> str:
    """sha256 over the bytes of every input tensor + the scalar settings."""
    import hashlib
    h = hashlib.sha256()
    for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp", "viewmatrix", "projmatrix",
              "projmatrix_raw", "campos", "bg", "grad_color", "grad_depth"):
        v = getattr(sc, k)
        if v is not None:
            h.update(k.encode())
            h.update(np.ascontiguousarray(v.detach().cpu().numpy()).tobytes())
    h.update(repr((sc.W, sc.H, float(np.float32(sc.tanfovx)), float(np.float32(sc.tanfovy)), sc.sh_degree, sc.scale_modifier)).encode())
    return h.hexdigest()


# BASELINE.json configs (workload names used by bench.py and the tests)
def config_scene(name: str, seed: int = 0) -> Scene:
    if name == "C1":    # 10k, 320x240, deg 0, CPU-runnable
        return make_scene(10_000, 320, 240, sh_degree=0, seed=seed, name="C1")
    if name == "C2":    # 100k, 640x480, deg 3, fwd+bwd with pose grads
        return make_scene(100_000, 640, 480, sh_degree=3, seed=seed, name="C2")
    if name == "C3":    # 500k, 640x480, deg 0 (SLAM default)
        return make_scene(500_000, 640, 480, sh_degree=0, seed=seed, name="C3")
    if name == "C3sh3":
        return make_scene(500_000, 640, 480, sh_degree=3, seed=seed, name="C3sh3")
    if name == "C4":    # 2M, 1280x960, deg 0
        return make_scene(2_000_000, 1280, 960, sh_degree=0, seed=seed, name="C4")
    raise KeyError(name)


def to_numpy(sc: Scene, dtype=np.float32) -> dict:
    out = {}
    for k, v in sc.__dict__.items():
        if isinstance(v, torch.Tensor):
            out[k] = np.ascontiguousarray(v.detach().cpu().numpy().astype(dtype))
        else:
            out[k] = v
    return out
