"""Pins the raw-parameter restatement (oracle.g4r_oracle.activate_raw / raw_chain_rule) to the reference's own prelude:
the torch calls of gaussian_splatting/scene/gaussian_model.py:100-128 (exp / sigmoid / F.normalize / cat) and the gradients
torch.autograd derives for them.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.g4r_oracle import activate_raw, raw_chain_rule  # noqa: E402


def _raw(P, M, scale_dim, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    return dict(opacity=r(P, 1) * 1.5, dc=r(P, 1, 3), rest=r(P, M - 1, 3) * 0.1, scaling=r(P, scale_dim) * 0.5 - 3.0,
                rotation=r(P, 4) * 2.0)


@pytest.mark.parametrize("M,scale_dim", [(1, 3), (4, 3), (16, 3), (1, 1), (9, 1)])
def test_activation_and_chain_rule_match_torch_autograd(M, scale_dim):
    P = 257
    raw = {k: v.clone().requires_grad_(True) for k, v in _raw(P, M, scale_dim, 7 + M).items()}
    # the reference prelude
    opac = torch.sigmoid(raw["opacity"])
    scal = torch.exp(raw["scaling"])
    scal3 = scal.repeat(1, 3) if scale_dim == 1 else scal
    rot = torch.nn.functional.normalize(raw["rotation"])
    shs = torch.cat((raw["dc"], raw["rest"]), dim=1)

    act = activate_raw(raw["opacity"].detach().numpy(), raw["dc"].detach().numpy(), raw["rest"].detach().numpy(),
                       raw["scaling"].detach().numpy(), raw["rotation"].detach().numpy(), dtype=np.float64)
    assert np.allclose(act["opacities"], opac.detach().numpy(), rtol=1e-13, atol=0)
    assert np.allclose(act["scales"], scal3.detach().numpy(), rtol=1e-13, atol=0)
    assert np.allclose(act["rotations"], rot.detach().numpy(), rtol=1e-13, atol=1e-15)
    assert np.array_equal(act["shs"], shs.detach().numpy())

    # random upstream gradients w.r.t. the activated tensors (what the rasterizer's backward delivers)
    g = torch.Generator().manual_seed(99)
    up = dict(o=torch.randn(P, 1, generator=g, dtype=torch.float64), s=torch.randn(P, 3, generator=g, dtype=torch.float64),
              q=torch.randn(P, 4, generator=g, dtype=torch.float64), sh=torch.randn(P, M, 3, generator=g, dtype=torch.float64))
    loss = (opac * up["o"]).sum() + (scal3 * up["s"]).sum() + (rot * up["q"]).sum() + (shs * up["sh"]).sum()
    loss.backward()
    mine = raw_chain_rule(act, dict(dL_dopacity=up["o"].numpy().reshape(-1), dL_dscales=up["s"].numpy(), dL_drots=up["q"].numpy(),
                                    dL_dshs=up["sh"].numpy()), scale_dim=scale_dim)
    assert np.allclose(mine["dL_dopacity_raw"], raw["opacity"].grad.numpy().reshape(-1), rtol=1e-11, atol=1e-14)
    assert np.allclose(mine["dL_dscaling_raw"], raw["scaling"].grad.numpy(), rtol=1e-11, atol=1e-14)
    assert np.allclose(mine["dL_drotation_raw"], raw["rotation"].grad.numpy(), rtol=1e-10, atol=1e-13)
    assert np.array_equal(mine["dL_dfeatures_dc"], raw["dc"].grad.numpy())
    assert np.array_equal(mine["dL_dfeatures_rest"], raw["rest"].grad.numpy())


def test_float32_activation_is_close_to_torch_float32():
    raw = _raw(1000, 4, 3, 3)
    act = activate_raw(*(raw[k].float().numpy() for k in ("opacity", "dc", "rest", "scaling", "rotation")))
    assert np.abs(act["opacities"] - torch.sigmoid(raw["opacity"].float()).numpy()).max() < 2e-7
    assert np.abs(act["scales"] / torch.exp(raw["scaling"].float()).numpy() - 1).max() < 3e-7
    assert np.abs(act["rotations"] - torch.nn.functional.normalize(raw["rotation"].float()).numpy()).max() < 2e-7
