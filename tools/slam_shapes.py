"""The two hot loops of 4DGS-SLAM restated around the rasterizer, for bench.py and the caller tests (test / bench
infrastructure; the product package never imports this).

* ``DuckGaussians`` -- the slice of ``GaussianModel`` that ``render()`` touches (gaussian_splatting/scene/gaussian_model.py:
  37-128: raw parameters ``_xyz/_features_dc/_features_rest/_scaling/_rotation/_opacity``, the activation properties, ``dygs``).
* ``DuckCamera``    -- the slice of ``Camera`` (utils/camera_utils.py:24-150): R, T, pose deltas, exposure, and the matrix
  properties recomputed per access exactly like the reference (two ``linalg.inv`` in ``getWorld2View2``,
  gaussian_splatting/utils/graphics_utils.py:33-46).
* ``render``        -- a restatement of the reference's ``render()`` (gaussian_splatting/gaussian_renderer/__init__.py:41-226)
  for the paths the SLAM loops use (SH colours, scale+rotation, optional static ``mask``, optional dynamic offsets
  ``dx/ds/dr``).  tests/test_reference_caller.py runs the UNMODIFIED reference file on the same ducks and compares.
* ``tracking_iteration`` / ``mapping_iteration`` -- the loop bodies of ``FrontEnd.tracking`` (utils/slam_frontend.py:411-448)
  and ``BackEnd.map`` (utils/slam_backend.py:357-771) with the RGB-D losses of utils/slam_utils.py:57-173,252-364.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch

PIPE = SimpleNamespace(convert_SHs_python=False, compute_cov3D_python=False)      # arguments.py:68-73


class DuckGaussians:
    def __init__(self, xyz, features_dc, features_rest, scaling, rotation, opacity, dygs=None, sh_degree=0, requires_grad=True):
        mk = (lambda t: torch.nn.Parameter(t.detach().clone().contiguous(), requires_grad=True)) if requires_grad else (lambda t: t.detach().clone())
        self._xyz, self._features_dc, self._features_rest = mk(xyz), mk(features_dc), mk(features_rest)
        self._scaling, self._rotation, self._opacity = mk(scaling), mk(rotation), mk(opacity)
        self.dygs = dygs if dygs is not None else torch.zeros(xyz.shape[0], dtype=torch.bool, device=xyz.device)
        self.active_sh_degree = sh_degree
        self.max_sh_degree = sh_degree
        self.scaling_activation = torch.exp
        self.opacity_activation = torch.sigmoid
        self.rotation_activation = torch.nn.functional.normalize

    @classmethod
    def from_scene(cls, sc, dygs=None, requires_grad=True):
        """Raw parameters whose activations reproduce the scene (tools/runners.raw_parameters)."""
        from tools import runners
        raw = runners.raw_parameters(sc, scale_dim=3, seed=0)
        return cls(raw["xyz"], raw["dc"], raw["rest"], raw["scaling"], raw["rotation"], raw["opacity"].view(-1, 1), dygs=dygs,
                   sh_degree=sc.sh_degree, requires_grad=requires_grad)

    # gaussian_model.py:100-128
    @property
    def get_scaling(self):
        return self.scaling_activation(self._scaling)

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return self.opacity_activation(self._opacity)

    def parameters(self):
        return [self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation]

    def optimizer(self):
        """Adam groups like GaussianModel.training_setup (gaussian_model.py:402-446; learning rates of configs/rgbd/tum)."""
        lr = dict(xyz=1.6e-4, f_dc=2.5e-3, f_rest=2.5e-3 / 20.0, opacity=0.05, scaling=1e-3, rotation=1e-3)
        groups = [{"params": [p], "lr": lr[n], "name": n} for n, p in zip(("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"), self.parameters())]
        return torch.optim.Adam(groups, lr=0.0, eps=1e-15)


def _world2view2(R, t):
    """getWorld2View2 (graphics_utils.py:33-46) with translate = 0, scale = 1: the matrix round-trips through its inverse."""
    Rt = torch.zeros((4, 4), device=R.device)
    Rt[:3, :3] = R
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = torch.linalg.inv(Rt)
    return torch.linalg.inv(C2W)


class DuckCamera:
    def __init__(self, uid, R, T, projection_matrix, FoVx, FoVy, H, W, image=None, depth=None, time=0.0):
        dev = R.device
        self.uid, self.device = uid, dev
        self.R, self.T = R.clone(), T.clone()
        self.projection_matrix = projection_matrix.to(dev)                 # = P^T (utils/camera_utils.py:82)
        self.FoVx, self.FoVy, self.image_height, self.image_width, self.time = FoVx, FoVy, H, W, time
        self.cam_rot_delta = torch.nn.Parameter(torch.zeros(3, device=dev))
        self.cam_trans_delta = torch.nn.Parameter(torch.zeros(3, device=dev))
        self.exposure_a = torch.nn.Parameter(torch.tensor([0.0], device=dev))
        self.exposure_b = torch.nn.Parameter(torch.tensor([0.0], device=dev))
        self.original_image, self.depth_gt = image, depth
        self.motion_mask, self.grad_mask = None, None

    @classmethod
    def from_scene(cls, sc, uid=0, image=None, depth=None):
        Rt = sc.viewmatrix.t()
        return cls(uid, Rt[:3, :3].contiguous(), Rt[:3, 3].contiguous(), sc.projmatrix_raw, 2 * math.atan(sc.tanfovx),
                   2 * math.atan(sc.tanfovy), sc.H, sc.W, image=image, depth=depth)

    # utils/camera_utils.py:124-148
    @property
    def world_view_transform(self):
        return _world2view2(self.R, self.T).transpose(0, 1)

    @property
    def full_proj_transform(self):
        return (self.world_view_transform.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0))).squeeze(0)

    @property
    def camera_center(self):
        return self.world_view_transform.inverse()[3, :3]

    def update_RT(self, R, t):
        self.R, self.T = R.to(self.device), t.to(self.device)


def perturbed(cam: DuckCamera, uid: int, rot=0.02, trans=0.05) -> DuckCamera:
    """A nearby keyframe: the same intrinsics, pose = SE3_exp(small tau) @ pose (deterministic in uid)."""
    g = torch.Generator().manual_seed(1000 + uid)
    tau = torch.cat([trans * (torch.rand(3, generator=g) - 0.5), rot * (torch.rand(3, generator=g) - 0.5)]).to(cam.device)
    T = torch.eye(4, device=cam.device)
    T[:3, :3], T[:3, 3] = cam.R, cam.T
    new = se3_exp(tau) @ T
    c = DuckCamera(uid, new[:3, :3].contiguous(), new[:3, 3].contiguous(), cam.projection_matrix, cam.FoVx, cam.FoVy,
                   cam.image_height, cam.image_width, image=cam.original_image, depth=cam.depth_gt, time=cam.time)
    return c


# ---- utils/pose_utils.py:4-97 restated ------------------------------------------------------------------------------------------
def _skew(v):
    z = torch.zeros((), device=v.device, dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def se3_exp(tau):
    rho, theta = tau[:3], tau[3:]
    W = _skew(theta)
    W2 = W @ W
    a = torch.norm(theta)
    I = torch.eye(3, device=tau.device, dtype=tau.dtype)
    if float(a) < 1e-5:
        R, V = I + W + 0.5 * W2, I + 0.5 * W + W2 / 6.0
    else:
        R = I + (torch.sin(a) / a) * W + ((1 - torch.cos(a)) / a**2) * W2
        V = I + W * ((1.0 - torch.cos(a)) / a**2) + W2 * ((a - torch.sin(a)) / a**3)
    T = torch.eye(4, device=tau.device, dtype=tau.dtype)
    T[:3, :3], T[:3, 3] = R, V @ rho
    return T


def update_pose(camera, converged_threshold=1e-4):
    tau = torch.cat([camera.cam_trans_delta, camera.cam_rot_delta], axis=0)
    T_w2c = torch.eye(4, device=tau.device)
    T_w2c[0:3, 0:3], T_w2c[0:3, 3] = camera.R, camera.T
    new_w2c = se3_exp(tau) @ T_w2c
    converged = tau.norm() < converged_threshold
    camera.update_RT(new_w2c[0:3, 0:3], new_w2c[0:3, 3])
    camera.cam_rot_delta.data.fill_(0)
    camera.cam_trans_delta.data.fill_(0)
    return converged


# ---- the caller ------------------------------------------------------------------------------------------------------------------
def render(dgr, viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, mask=None, dx=None, ds=None, dr=None):
    """gaussian_renderer/__init__.py:41-226 for the paths the SLAM loops take; `dgr` = the rasterizer package to call."""
    if pc.get_xyz.shape[0] == 0:
        return None
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device=pc.get_xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:       # noqa: BLE001
        pass
    raster_settings = dgr.GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = dgr.GaussianRasterizer(raster_settings=raster_settings)
    means3D, means2D, opacity = pc.get_xyz, screenspace_points, pc.get_opacity
    scales = pc.get_scaling.repeat(1, 3) if pc.get_scaling.shape[-1] == 1 else pc.get_scaling
    rotations = pc.get_rotation
    shs = pc.get_features
    if dx is not None and ds is not None and dr is not None:                 # :159-174
        dxyz = torch.zeros_like(means3D)
        dxyz[pc.dygs] = dx
        means3D = pc.get_xyz + dxyz
        dscale = torch.zeros_like(scales)
        dscale[pc.dygs] = ds
        scales = scales + dscale
        drot = torch.zeros_like(rotations)
        drot[pc.dygs] = dr
        rotations = pc.get_rotation + drot
    if mask is not None:                                                     # :180-191
        rendered_image, radii, depth, opacity, n_touched = rasterizer(
            means3D=means3D[mask], means2D=means2D[mask], shs=shs[mask], colors_precomp=None, opacities=opacity[mask],
            scales=scales[mask], rotations=rotations[mask], cov3D_precomp=None, theta=viewpoint_camera.cam_rot_delta,
            rho=viewpoint_camera.cam_trans_delta)
    else:
        rendered_image, radii, depth, opacity, n_touched = rasterizer(
            means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None, opacities=opacity, scales=scales,
            rotations=rotations, cov3D_precomp=None, theta=viewpoint_camera.cam_rot_delta, rho=viewpoint_camera.cam_trans_delta)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "opacity": opacity, "n_touched": n_touched}


# ---- losses (utils/slam_utils.py) -----------------------------------------------------------------------------------------------
RGB_BOUNDARY_THRESHOLD = 0.01          # configs/rgbd/tum/base_config.yaml
ALPHA = 0.95


def loss_tracking_rgbd(image, depth, opacity, viewpoint, alpha=ALPHA):
    """get_loss_tracking -> get_loss_tracking_rgbd (slam_utils.py:57-173) without the plotting side effects."""
    image_ab = torch.exp(viewpoint.exposure_a) * image + viewpoint.exposure_b
    gt_image, gt_depth = viewpoint.original_image, viewpoint.depth_gt
    rgb_pixel_mask = (gt_image.sum(dim=0) > RGB_BOUNDARY_THRESHOLD).view(1, *gt_image.shape[1:])
    if viewpoint.grad_mask is not None:
        rgb_pixel_mask = rgb_pixel_mask * viewpoint.grad_mask
    if viewpoint.motion_mask is not None and viewpoint.uid > 0:
        rgb_pixel_mask = viewpoint.motion_mask.view(*rgb_pixel_mask.shape) * rgb_pixel_mask
    l1_rgb = (opacity * torch.abs(image_ab * rgb_pixel_mask - gt_image * rgb_pixel_mask)).mean()
    depth_pixel_mask = (gt_depth > 0.01).view(*depth.shape)
    depth_pixel_mask = depth_pixel_mask * (gt_depth < 1000.).view(*depth.shape)
    depth_mask = depth_pixel_mask * (opacity > 0.95).view(*depth.shape)
    if viewpoint.motion_mask is not None and viewpoint.uid > 0:
        depth_mask = viewpoint.motion_mask.view(*depth.shape) * depth_mask
    l1_depth = torch.abs(depth * depth_mask - gt_depth * depth_mask)
    return alpha * l1_rgb + (1 - alpha) * l1_depth.mean()


def loss_mapping_rgbd(image, depth, viewpoint, alpha=ALPHA):
    """get_loss_mapping -> get_loss_mapping_rgbd (slam_utils.py:252-364), static branch."""
    image_ab = torch.exp(viewpoint.exposure_a) * image + viewpoint.exposure_b
    gt_image, gt_depth = viewpoint.original_image, viewpoint.depth_gt
    rgb_pixel_mask = (gt_image.sum(dim=0) > RGB_BOUNDARY_THRESHOLD).view(*depth.shape)
    depth_pixel_mask = (gt_depth > 0.01).view(*depth.shape)
    depth_pixel_mask = depth_pixel_mask * (gt_depth < 10000.).view(*depth.shape)
    if viewpoint.motion_mask is not None:
        rgb_pixel_mask = viewpoint.motion_mask.view(*depth.shape) * rgb_pixel_mask
        depth_pixel_mask = viewpoint.motion_mask.view(*depth.shape) * depth_pixel_mask
    l1_rgb = torch.abs(image_ab * rgb_pixel_mask - gt_image * rgb_pixel_mask)
    l1_depth = torch.abs(depth * depth_pixel_mask - gt_depth * depth_pixel_mask)
    return alpha * l1_rgb.mean() + (1 - alpha) * l1_depth.mean()


def isotropic_loss(scaling):
    """utils/slam_backend.py:645-647: |s - mean(s)| regulariser added to the mapping loss."""
    return torch.abs(scaling - scaling.mean(dim=1).view(-1, 1)).mean()


# ---- loop bodies ----------------------------------------------------------------------------------------------------------------
def pose_optimizer(viewpoint):
    """utils/slam_frontend.py:355-385."""
    return torch.optim.Adam([
        {"params": [viewpoint.cam_rot_delta], "lr": 0.003, "name": "rot"},
        {"params": [viewpoint.cam_trans_delta], "lr": 0.001, "name": "trans"},
        {"params": [viewpoint.exposure_a], "lr": 0.01, "name": "exposure_a"},
        {"params": [viewpoint.exposure_b], "lr": 0.01, "name": "exposure_b"}])


def tracking_iteration(dgr, render_fn, viewpoint, gaussians, bg, opt, gauss_opt=None):
    """One pass of the loop at utils/slam_frontend.py:411-448: masked static render, RGB-D tracking loss, backward, pose Adam
    step, zero the Gaussian gradients (they are computed and thrown away by the reference), SE(3) pose update."""
    pkg = render_fn(viewpoint, gaussians, PIPE, bg, mask=(gaussians.dygs == False))      # noqa: E712  (the reference's spelling)
    loss = loss_tracking_rgbd(pkg["render"], pkg["depth"], pkg["opacity"], viewpoint)
    loss.backward()
    with torch.no_grad():
        opt.step()
        opt.zero_grad()
        if gauss_opt is not None:
            gauss_opt.zero_grad(set_to_none=True)
        update_pose(viewpoint)
    return loss


def mapping_iteration(dgr, render_fn, viewpoints, gaussians, bg, gauss_opt, pose_opts):
    """One iteration of BackEnd.map (utils/slam_backend.py:357-771): every window keyframe (+ random older ones) is rendered
    and its loss accumulated, ONE backward(retain_graph=True) runs through all the rasterizer nodes, then the Gaussian and
    pose Adam steps."""
    loss_mapping = 0
    for vp in viewpoints:
        pkg = render_fn(vp, gaussians, PIPE, bg)
        loss_mapping = loss_mapping + loss_mapping_rgbd(pkg["render"], pkg["depth"], vp)
    loss_mapping = loss_mapping + 10 * isotropic_loss(gaussians.get_scaling)
    loss_mapping.backward(retain_graph=True)
    with torch.no_grad():
        gauss_opt.step()
        gauss_opt.zero_grad(set_to_none=True)
        for vp, po in zip(viewpoints, pose_opts):
            po.step()
            po.zero_grad(set_to_none=True)
            update_pose(vp)
    return loss_mapping


# ---- the same loop bodies on the opt-in fused API ------------------------------------------------------
def camera_matrices_once(vp):
    """The three camera tensors of the raster settings from ONE evaluation of world_view_transform (the reference's Camera
    properties recompute it -- two linalg.inv each time -- for viewmatrix, full_proj_transform and camera_center,
    utils/camera_utils.py:124-148: seven 4x4 inversions per render)."""
    view = vp.world_view_transform
    proj = view.unsqueeze(0).bmm(vp.projection_matrix.unsqueeze(0)).squeeze(0)
    return view, proj, view.inverse()[3, :3]


def render_fused(dgr, vp, params, bg, sh_degree=0, mask=None, with_screenspace_grad=True):
    """render() on FusedGaussianRasterizer: raw parameters in (activations, cat, mask gather inside the kernels).
    `params` = (xyz, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw)."""
    view, proj, campos = camera_matrices_once(vp)
    rs = dgr.GaussianRasterizationSettings(
        image_height=int(vp.image_height), image_width=int(vp.image_width), tanfovx=math.tan(vp.FoVx * 0.5), tanfovy=math.tan(vp.FoVy * 0.5),
        bg=bg, scale_modifier=1.0, viewmatrix=view, projmatrix=proj, projmatrix_raw=vp.projection_matrix, sh_degree=sh_degree, campos=campos,
        prefiltered=False, debug=False)
    xyz, f_dc, f_rest, opacity, scaling, rotation = params
    screenspace_points = torch.zeros_like(xyz, requires_grad=with_screenspace_grad)
    image, radii, depth, opacity_img, n_touched = dgr.FusedGaussianRasterizer(rs)(
        xyz=xyz, means2D=screenspace_points, features_dc=f_dc, features_rest=f_rest, opacity_raw=opacity, scaling_raw=scaling,
        rotation_raw=rotation, theta=vp.cam_rot_delta, rho=vp.cam_trans_delta, mask=mask)
    return {"render": image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii, "depth": depth,
            "opacity": opacity_img, "n_touched": n_touched}


def tracking_iteration_fused(dgr, viewpoint, frozen_params, static_mask, bg, opt, sh_degree=0):
    """utils/slam_frontend.py:411-448 on the fused API: the map is frozen during tracking, so the (detached) raw parameters and
    the static mask are prepared once per frame; per iteration: one masked render, one loss kernel, a pose-only backward (no
    per-Gaussian gradient is computed, let alone zeroed afterwards), pose Adam step, SE(3) update."""
    from diff_gaussian_rasterization.losses import slam_loss
    pkg = render_fused(dgr, viewpoint, frozen_params, bg, sh_degree=sh_degree, mask=static_mask, with_screenspace_grad=False)
    loss = slam_loss("tracking", pkg["render"], pkg["depth"], viewpoint.original_image, viewpoint.depth_gt, opacity=pkg["opacity"],
                     exposure_a=viewpoint.exposure_a, exposure_b=viewpoint.exposure_b,
                     motion_mask=viewpoint.motion_mask if viewpoint.uid > 0 else None, grad_mask=viewpoint.grad_mask,
                     alpha=ALPHA, rgb_boundary_threshold=RGB_BOUNDARY_THRESHOLD)
    loss.backward()
    with torch.no_grad():
        opt.step()
        opt.zero_grad()
        update_pose(viewpoint)
    return loss


def mapping_iteration_fused(dgr, viewpoints, gaussians, bg, gauss_opt, pose_opts):
    """utils/slam_backend.py:357-771 on the fused API: per view one fused render (no torch prelude) and one loss kernel."""
    from diff_gaussian_rasterization.losses import slam_loss
    loss_mapping = 0
    params = tuple(gaussians.parameters()[i] for i in (0, 1, 2, 3, 4, 5))
    for vp in viewpoints:
        pkg = render_fused(dgr, vp, params, bg, sh_degree=gaussians.active_sh_degree)
        loss_mapping = loss_mapping + slam_loss("mapping", pkg["render"], pkg["depth"], vp.original_image, vp.depth_gt,
                                                exposure_a=vp.exposure_a, exposure_b=vp.exposure_b, motion_mask=vp.motion_mask, alpha=ALPHA,
                                                rgb_boundary_threshold=RGB_BOUNDARY_THRESHOLD)
    loss_mapping = loss_mapping + 10 * isotropic_loss(gaussians.get_scaling)
    loss_mapping.backward(retain_graph=True)
    with torch.no_grad():
        gauss_opt.step()
        gauss_opt.zero_grad(set_to_none=True)
        for vp, po in zip(viewpoints, pose_opts):
            po.step()
            po.zero_grad(set_to_none=True)
            update_pose(vp)
    return loss_mapping
