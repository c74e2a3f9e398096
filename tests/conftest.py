import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "4dgs-slam_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # Build the CUDA library and the oracle in-tree BEFORE collection imports the package (nvcc / gcc cross-compile
    # without a GPU; both are no-ops when the binaries are newer than their sources, e.g. on the GPU box).
    import __graft_entry__ as ge
    ge.build_library()
    try:
        ge.build_host()          # optional C++ host side of the standard op; the Python path is used when it is absent
    except Exception as exc:     # noqa: BLE001
        sys.stderr.write(f"build_host failed ({exc!r}); tests run on the Python host path\n")
    ge.build_oracle()


@pytest.fixture(scope="session")
def device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
