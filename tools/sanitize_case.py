"""Small forward+backward cases for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools import runners
from tools.scenes import make_scene

dev = torch.device("cuda:0")
for kw in (dict(P=700, W=100, H=75, sh_degree=3, seed=1), dict(P=500, W=64, H=48, sh_degree=0, seed=2, colors_precomp=True, cov3D_precomp=True),
           dict(P=6000, W=48, H=32, sh_degree=0, seed=3, px_min=4.0, px_max=30.0)):
    out = runners.run_g4r(make_scene(**kw).to(dev))
    torch.cuda.synchronize()
    print(kw, "N", out["num_rendered"])
print("sanitize cases done")
