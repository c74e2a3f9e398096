"""Run the reference's caller -- ``gaussian_splatting/gaussian_renderer/__init__.py`` (``render`` / ``render_flow``), the file
BASELINE.json's north_star says must call the rasterizer unchanged -- UNMODIFIED on top of a rasterizer package of our choice.

The reference tree does not exist on the GPU box, so ``install()`` (called by ``__graft_entry__.build()`` in the container
that has /root/reference) copies the two files the caller needs, byte for byte, next to the reference rasterizer build in the
git-ignored ``baseline/_ref`` (which travels with gpurun, like the rasterizer build itself).  Nothing is copied into the
tracked tree.  ``load_renderer(dgr)`` executes the copied file with

* ``diff_gaussian_rasterization`` resolved to `dgr` (this repo's drop-in, or the reference build loaded by tools/refload),
* ``gaussian_splatting.scene.gaussian_model`` stubbed (the caller only uses ``GaussianModel`` as a type annotation; the real
  module pulls open3d / plyfile / simple_knn, which are not installed here),

and returns the module object.  Test / bench infrastructure only."""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"
CALLER_DIR = os.path.join(ROOT, "baseline", "_ref", "caller")
FILES = ("gaussian_splatting/gaussian_renderer/__init__.py", "gaussian_splatting/utils/sh_utils.py")


def install(force: bool = False) -> bool:
    if available() and not force:
        return True
    if not os.path.isdir(REF_ROOT):
        return False
    for rel in FILES:
        dst = os.path.join(CALLER_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(CALLER_DIR, rel)) for rel in FILES)


def _pkg(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = []          # a package
    return m


def load_renderer(dgr, tag: str):
    """The reference's gaussian_renderer module, executed against rasterizer package `dgr`; `tag` names the module object."""
    if not available():
        raise ImportError("baseline/_ref/caller is absent: run __graft_entry__.build() in a container that has /root/reference")
    saved = {k: sys.modules.get(k) for k in ("diff_gaussian_rasterization", "gaussian_splatting", "gaussian_splatting.scene",
                                             "gaussian_splatting.scene.gaussian_model", "gaussian_splatting.utils",
                                             "gaussian_splatting.utils.sh_utils")}
    try:
        sys.modules["diff_gaussian_rasterization"] = dgr
        sys.modules["gaussian_splatting"] = _pkg("gaussian_splatting")
        sys.modules["gaussian_splatting.scene"] = _pkg("gaussian_splatting.scene")
        gm = types.ModuleType("gaussian_splatting.scene.gaussian_model")
        gm.GaussianModel = type("GaussianModel", (), {})
        sys.modules["gaussian_splatting.scene.gaussian_model"] = gm
        sys.modules["gaussian_splatting.utils"] = _pkg("gaussian_splatting.utils")
        spec = importlib.util.spec_from_file_location("gaussian_splatting.utils.sh_utils", os.path.join(CALLER_DIR, FILES[1]))
        sh = importlib.util.module_from_spec(spec)
        sys.modules["gaussian_splatting.utils.sh_utils"] = sh
        spec.loader.exec_module(sh)
        spec = importlib.util.spec_from_file_location(f"ref_gaussian_renderer_{tag}", os.path.join(CALLER_DIR, FILES[0]))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.GaussianRasterizer is dgr.GaussianRasterizer
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
