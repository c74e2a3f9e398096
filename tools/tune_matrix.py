"""Kernel-variant experiment (run under gpurun, 1 GPU): each G4R_TUNE_* combination runs in its own process (the switches
are read once per process), on a few scene types; reports per-kernel device time (library stage profile), eager and
CUDA-graph fwd+bwd time, and checks that every variant reproduces the first one's outputs (images / integer outputs by
hash, gradients by relative L2)."""
import ctypes
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "4dgs-slam_b200"))

VARIANTS = {
    "default": {},
    "sorted splat stream + TMA bulk staging (STREAM=1)": {"STREAM": 1},
}
if os.environ.get("G4R_VARIANTS"):          # e.g. G4R_VARIANTS='{"x": {"LPT": 0}}'
    VARIANTS = {"default": {}, **json.loads(os.environ["G4R_VARIANTS"])}
# Round-1 history (profiles/r01_v7_tune_matrix.json) also covered shapes that were measured and dropped from the source:
# forward capped at 56 / 48 registers (9 / 10 CTAs per SM), backward CTA per tile with 64 / 128 staged splats, backward CTA
# per half tile with 128 staged.


SCENE_FILE = "/tmp/g4r_tune_scenes.pt"


def scenes():
    """Built once by the parent and shared through a file: the CPU generator is not guaranteed to round identically in
    every process, and the variants are compared by output hash."""
    import torch
    if os.path.exists(SCENE_FILE):
        return torch.load(SCENE_FILE, weights_only=False)
    from tools.perf_matrix import clustered
    from tools.scenes import config_scene, make_scene
    return {
        "C3": config_scene("C3"),
        "clustered200k": clustered(200_000, 640, 480, 4),
        "bigsplats100k": make_scene(100_000, 640, 480, sh_degree=0, seed=3, px_min=3.0, px_max=25.0),
        "C2": config_scene("C2"),
        "C4": config_scene("C4"),
    }


def worker():
    import torch
    import diff_gaussian_rasterization as dgr
    from tools import runners
    from tools.perf_matrix import graph_ms, timeit
    lib = dgr._lib
    lib.g4r_profile_stage_name.restype = ctypes.c_char_p
    dev = torch.device("cuda:0")
    out = {}
    for name, sc_cpu in scenes().items():
        sc = sc_cpu.to(dev)
        r = runners.run_public_api(sc, dgr)
        torch.cuda.synchronize()
        row = {}
        for k in ("color", "depth", "opacity", "radii", "n_touched"):
            row["sha_" + k] = hashlib.sha1(r[k].detach().cpu().numpy().tobytes()).hexdigest()[:16]
        gpath = f"/tmp/g4r_tune_grads_{name}.pt"
        base_g = torch.load(gpath) if os.path.exists(gpath) else None
        save_g = {}
        for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dshs", "dL_dscales", "dL_drots", "dL_dtau"):
            if r.get(k) is not None:
                g = r[k].detach().double()
                row["g_" + k] = [float(g.norm()), float(g.sum())]
                save_g[k] = g.cpu()
                if base_g is not None:       # relative L2 distance to the first (default) variant's gradient
                    b = base_g[k].to(g.device)
                    row["gl2_" + k] = float((g - b).norm() / (b.norm() + 1e-30))
        if base_g is None:
            torch.save(save_g, gpath)
        lib.g4r_profile_enable(1)
        row["eager_ms"] = timeit(lambda: runners.run_public_api(sc, dgr), warmup=3, iters=20)
        n_st = lib.g4r_profile_stage_count()
        ms_arr = (ctypes.c_double * n_st)()
        cnt_arr = (ctypes.c_int64 * n_st)()
        lib.g4r_profile_read(ms_arr, cnt_arr, 1)
        lib.g4r_profile_enable(0)
        row["stage_us"] = {lib.g4r_profile_stage_name(i).decode(): round(1e3 * ms_arr[i] / max(1, cnt_arr[i]), 1) for i in range(n_st)}
        try:
            row["graph_ms"], row["graph_ok"] = graph_ms(sc, dgr, iters=20)
        except Exception as exc:     # noqa: BLE001
            row["graph_ms"] = f"failed: {exc!r}"[:200]
        out[name] = row
    print("TUNE_RESULT " + json.dumps(out), flush=True)


def main():
    import torch
    import glob
    for f in [SCENE_FILE] + glob.glob("/tmp/g4r_tune_grads_*.pt"):
        if os.path.exists(f):
            os.remove(f)
    torch.save(scenes(), SCENE_FILE)
    results = {}
    for vname, env_add in VARIANTS.items():
        env = dict(os.environ)
        for k, v in env_add.items():
            env["G4R_TUNE_" + k] = str(v)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True, timeout=240)
        except subprocess.TimeoutExpired as exc:      # a hung kernel: the child is killed, the other variants still run
            results[vname] = {"error": "timeout: " + str(exc.stdout)[-800:]}
            print(vname, "TIMEOUT", flush=True)
            continue
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("TUNE_RESULT ")]
        if p.returncode != 0 or not line:
            results[vname] = {"error": p.stdout[-1500:]}
            print(vname, "FAILED", p.stdout[-1500:], flush=True)
            continue
        results[vname] = json.loads(line[0][len("TUNE_RESULT "):])
    base = results.get("default", {})
    for vname, res in results.items():
        if "error" in res:
            continue
        for sname, row in res.items():
            b = base.get(sname, {})
            same = all(row[k] == b.get(k) for k in row if k.startswith("sha_"))
            gerr = 0.0
            for k in row:
                if k.startswith("g_") and k in b:
                    gerr = max(gerr, abs(row[k][0] - b[k][0]) / (abs(b[k][0]) + 1e-30))
            row["outputs_identical_to_default"] = same
            row["grad_norm_rel_diff"] = gerr
            gl2 = max([row[k] for k in row if k.startswith("gl2_")] + [0.0])
            row["grad_l2_rel_to_default_max"] = gl2
            gerr = max(gerr, gl2)
            st = row["stage_us"]
            print(f"{vname:38s} {sname:14s} eager {row['eager_ms']:.3f} graph {row['graph_ms'] if isinstance(row['graph_ms'], str) else round(row['graph_ms'], 3)}"
                  f" | fwd {st.get('composite_forward')} bwd {st.get('composite_backward')} sort {st.get('tile_sort')} scan {st.get('tile_scan')}"
                  f" | same={same} gerr={gerr:.1e}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tune_matrix.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    if "--worker" in sys.argv:
        worker()
    else:
        main()
