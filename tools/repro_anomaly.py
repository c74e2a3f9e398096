"""Hunt for a rare forward-output mismatch seen once across processes (run under gpurun, 1 GPU).

Every worker process renders C3 through the autograd API several times.  Per step it snapshots colour/depth/opacity right
after the forward, runs the backward, and then checks (a) the live output tensors still equal their snapshot (would catch a
backward kernel writing out of bounds), (b) the snapshot equals a forward-only render with the saved state, and (c) the
snapshot equals the first process's snapshot (kept in /tmp on the box).  Any difference is localised (pixels, tile, values)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "4dgs-slam_b200"))
GOLD = "/tmp/g4r_repro_gold.pt"


def where(a, b, W, gx, limit=12):
    neq = (a.view(-1).view(dtype=__import__("torch").int32) != b.view(-1).view(dtype=__import__("torch").int32))
    idx = neq.nonzero().reshape(-1)
    plane = a.shape[-1] * a.shape[-2]
    out = []
    for i in idx[:limit].tolist():
        c, pix = divmod(i, plane)
        y, x = divmod(pix, W)
        out.append({"c": c, "x": x, "y": y, "tile": (y // 16) * gx + x // 16, "a": float(a.view(-1)[i]), "b": float(b.view(-1)[i])})
    return {"count": int(idx.numel()), "first": out}


def worker(steps):
    import torch
    import diff_gaussian_rasterization as dgr
    from tools import runners
    from tools.scenes import config_scene
    dev = torch.device("cuda:0")
    sc = config_scene("C3").to(dev)
    rs = runners.settings_for(sc, dgr)
    W, gx = sc.W, (sc.W + 15) // 16
    gold = torch.load(GOLD) if os.path.exists(GOLD) else None
    rep = {"steps": steps, "events": []}
    for step in range(steps):
        leaf = {k: getattr(sc, k).detach().clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        m2d = torch.zeros_like(leaf["means3D"], requires_grad=True)
        theta = torch.zeros(3, device=dev, requires_grad=True)
        rho = torch.zeros(3, device=dev, requires_grad=True)
        color, radii, depth, opacity, n_touched = dgr.GaussianRasterizer(rs)(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf["shs"], colors_precomp=None,
            scales=leaf["scales"], rotations=leaf["rotations"], cov3D_precomp=None, theta=theta, rho=rho)
        snap = {"color": color.detach().clone(), "depth": depth.detach().clone(), "opacity": opacity.detach().clone()}
        loss = (color * sc.grad_color).sum() + (depth * sc.grad_depth).sum()
        loss.backward()
        torch.cuda.synchronize()
        live = {"color": color.detach(), "depth": depth.detach(), "opacity": opacity.detach()}
        out, info = dgr.rasterize_gaussians_with_state(rs, sc.means3D, sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
        again = {"color": out[0], "depth": out[2], "opacity": out[3]}
        ev = {}
        for k in snap:
            if not torch.equal(snap[k], live[k]):
                ev["after_backward_" + k] = where(snap[k], live[k], W, gx)
            if not torch.equal(snap[k], again[k]):
                ev["vs_forward_only_" + k] = where(snap[k], again[k], W, gx)
            if gold is not None and not torch.equal(snap[k].cpu(), gold[k]):
                ev["vs_first_process_" + k] = where(snap[k].cpu(), gold[k], W, gx)
        if gold is None and step == 0:
            torch.save({k: v.cpu() for k, v in snap.items()}, GOLD)
            gold = {k: v.cpu() for k, v in snap.items()}
        if ev:
            ev["step"] = step
            if len(rep["events"]) < 5:
                rep["events"].append(ev)
            rep["bad_steps"] = rep.get("bad_steps", 0) + 1
    print("REPRO_RESULT " + json.dumps(rep), flush=True)


def main():
    if os.path.exists(GOLD):
        os.remove(GOLD)
    n_proc = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    out = {}
    for p in range(n_proc):
        env = dict(os.environ)
        if p % 2 == 1:
            env["G4R_TUNE_LPT"] = "0"
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", "6"], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("REPRO_RESULT ")]
        res = json.loads(line[0][len("REPRO_RESULT "):]) if line else {"error": r.stdout[-1200:]}
        out[str(p)] = res
        print(p, "LPT=0" if p % 2 else "LPT=1", json.dumps(res)[:2500], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "repro_anomaly.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if "--worker" in sys.argv:
        worker(int(sys.argv[sys.argv.index("--worker") + 1]))
    else:
        main()
