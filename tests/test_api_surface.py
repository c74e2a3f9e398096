"""Drop-in surface: same names, field order, argument checks and error messages as the reference package
(DGR/diff_gaussian_rasterization/__init__.py:173-244); no CPU fallback."""
import inspect

import pytest
import torch

import diff_gaussian_rasterization as dgr


def test_exports():
    assert hasattr(dgr, "GaussianRasterizationSettings") and hasattr(dgr, "GaussianRasterizer")
    assert hasattr(dgr, "rasterize_gaussians") and hasattr(dgr, "_RasterizeGaussians")


def test_settings_fields_match_reference_order():
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "projmatrix_raw", "sh_degree", "campos", "prefiltered", "debug")


def test_forward_signature_matches_reference():
    sig = inspect.signature(dgr.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations",
                                    "cov3D_precomp", "theta", "rho"]
    for k in ("shs", "colors_precomp", "scales", "rotations", "cov3D_precomp", "theta", "rho"):
        assert sig.parameters[k].default is None
    assert list(inspect.signature(dgr.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp", "theta", "rho",
        "raster_settings"]


def _settings():
    e = torch.eye(4)
    return dgr.GaussianRasterizationSettings(image_height=32, image_width=32, tanfovx=1.0, tanfovy=1.0, bg=torch.ones(3),
                                             scale_modifier=1.0, viewmatrix=e, projmatrix=e, projmatrix_raw=e, sh_degree=0,
                                             campos=torch.zeros(3), prefiltered=False, debug=False)


def test_argument_combination_errors_match_reference_messages():
    r = dgr.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.ones(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.ones(4, 1), shs=torch.ones(4, 1, 3), colors_precomp=torch.ones(4, 3), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(4, 1), shs=torch.ones(4, 1, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(4, 1), shs=torch.ones(4, 1, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4),
          cov3D_precomp=torch.ones(4, 6))


def test_no_cpu_fallback():
    r = dgr.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(m, m, torch.ones(4, 1), shs=torch.ones(4, 1, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        r.markVisible(m)


def test_means3d_shape_error():
    r = dgr.GaussianRasterizer(_settings())
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        r(torch.zeros(4, 2), torch.zeros(4, 2), torch.ones(4, 1), shs=torch.ones(4, 1, 3), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4))


def test_product_never_imports_the_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "4dgs-slam_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "g4r_oracle" not in text.replace("oracle/g4r_oracle.c follows", "") or f == "project.cu", (dp, f)
                assert "import oracle" not in text and "from oracle" not in text


def test_fused_rasterizer_surface_and_argument_checks():
    """The opt-in raw-parameter module: exported, constructed like GaussianRasterizer, rejects CPU tensors
    loudly (no CPU path) and malformed raw tensors with a message."""
    import inspect

    import pytest
    import torch
    import diff_gaussian_rasterization as dgr
    assert "FusedGaussianRasterizer" in dgr.__all__ and "rasterize_gaussians_raw" in dgr.__all__
    sig = inspect.signature(dgr.FusedGaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["xyz", "means2D", "features_dc", "features_rest", "opacity_raw", "scaling_raw", "rotation_raw",
                                        "theta", "rho", "mask", "dx", "ds", "dr", "dyn_slot"]
    assert "dynamic_slots" in dgr.__all__
    slots = dgr.dynamic_slots(torch.tensor([False, True, True, False, True]))
    assert slots.tolist() == [-1, 0, 1, -1, 2] and slots.dtype is torch.int32
    rs = dgr.GaussianRasterizationSettings(image_height=8, image_width=8, tanfovx=1.0, tanfovy=1.0, bg=torch.ones(3), scale_modifier=1.0,
                                           viewmatrix=torch.eye(4), projmatrix=torch.eye(4), projmatrix_raw=torch.eye(4), sh_degree=0,
                                           campos=torch.zeros(3), prefiltered=False, debug=False)
    P = 5
    args = dict(xyz=torch.zeros(P, 3), means2D=torch.zeros(P, 3), features_dc=torch.zeros(P, 1, 3), features_rest=torch.zeros(P, 0, 3),
                opacity_raw=torch.zeros(P, 1), scaling_raw=torch.zeros(P, 3), rotation_raw=torch.ones(P, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dgr.FusedGaussianRasterizer(rs)(**args)
    with pytest.raises(RuntimeError, match="scaling_raw"):
        dgr.FusedGaussianRasterizer(rs)(**dict(args, scaling_raw=torch.zeros(P, 2)))
    with pytest.raises(RuntimeError, match="features_dc"):
        dgr.FusedGaussianRasterizer(rs)(**dict(args, features_dc=torch.zeros(P, 3)))


def test_debug_mode_writes_the_argument_snapshot_on_failure(tmp_path, monkeypatch):
    """debug=True: like the reference (DGR/diff_gaussian_rasterization/__init__.py:90-97) the arguments are copied to the host
    before the call and written to snapshot_fw.dump when the call raises.  (A CPU tensor is the failure that needs no GPU.)"""
    import pytest
    import torch
    import diff_gaussian_rasterization as dgr
    monkeypatch.chdir(tmp_path)
    eye = torch.eye(4)
    rs = dgr.GaussianRasterizationSettings(image_height=32, image_width=32, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3), scale_modifier=1.0,
                                           viewmatrix=eye, projmatrix=eye, projmatrix_raw=eye, sh_degree=0, campos=torch.zeros(3),
                                           prefiltered=False, debug=True)
    P = 5
    with pytest.raises(RuntimeError):
        dgr.GaussianRasterizer(rs)(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), opacities=torch.ones(P, 1), shs=torch.zeros(P, 1, 3),
                                   scales=torch.ones(P, 3), rotations=torch.ones(P, 4))
    snap = torch.load(tmp_path / "snapshot_fw.dump", weights_only=False)
    assert isinstance(snap, tuple) and snap[0].shape == (P, 3)


def test_simple_knn_drop_in_surface():
    """`from simple_knn._C import distCUDA2` (gaussian_splatting/scene/gaussian_model.py:18) resolves to this repo's module when
    4dgs-slam_b200 is on the path; one positional tensor argument like submodules/simple-knn/ext.cpp; no oracle on the product path."""
    import inspect

    import simple_knn._C as knn
    import os
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(dgr.__file__)))
    assert os.path.abspath(knn.__file__).startswith(pkg)
    assert list(inspect.signature(knn.distCUDA2).parameters) == ["points"]
    src = open(knn.__file__).read()
    assert "oracle" not in src


def test_control_node_warp_surface():
    """The opt-in warp takes the tensors ControlNodeWarp owns (utils/time_utils.py:824-828: nodes, _node_radius, _node_weight) and the
    node MLP's output dict, and returns the keys its forward returns (:1248-1275)."""
    import inspect

    from diff_gaussian_rasterization import deform
    params = list(inspect.signature(deform.control_node_warp).parameters)
    assert params[:6] == ["x", "nodes", "log_radius", "weight_logit", "node_attrs", "motion_mask"]
    d = inspect.signature(deform.control_node_warp).parameters
    assert d["K"].default == 3 and d["d_rot_as_res"].default is True and d["local_frame"].default is True     # arguments.py:107,118,119
    assert "oracle" not in open(deform.__file__).read()
