"""Forward determinism soak (run under gpurun, 1 GPU): the forward has no order-dependent arithmetic (unique sort keys,
sequential per-pixel compositing, integer atomics), so every repetition must reproduce the first one bit for bit.  Repeats
the forward of a scene many times, in several fresh processes, and reports WHICH state differs and where if one does."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "4dgs-slam_b200"))


def worker(workload, reps):
    import torch
    import diff_gaussian_rasterization as dgr
    from tools import runners
    from tools.scenes import config_scene
    dev = torch.device("cuda:0")
    sc = config_scene(workload).to(dev)
    rs = runners.settings_for(sc, dgr)

    def run():
        out, info = dgr.rasterize_gaussians_with_state(rs, sc.means3D, sc.opacities, shs=sc.shs, colors_precomp=sc.colors_precomp,
                                                       scales=sc.scales, rotations=sc.rotations, cov3D_precomp=sc.cov3D_precomp)
        d = dict(color=out[0], radii=out[1], depth=out[2], opacity=out[3], n_touched=out[4])
        d.update({k: v for k, v in info.items() if torch.is_tensor(v)})
        return {k: v.clone() for k, v in d.items()}

    first = run()
    torch.cuda.synchronize()
    import hashlib
    report = {"reps": reps, "mismatching_reps": 0, "events": [],
              "sha_first": {k: hashlib.sha1(v.cpu().numpy().tobytes()).hexdigest()[:12] for k, v in first.items()}}
    W = sc.W
    gx = (W + 15) // 16
    for r in range(reps):
        cur = run()
        bad = {}
        for k, v in cur.items():
            a = first[k]
            neq = (v.view(torch.int32) != a.view(torch.int32)) if v.dtype == torch.float32 else (v != a)
            n = int(neq.sum())
            if n:
                idx = neq.reshape(-1).nonzero().reshape(-1)[:8].tolist()
                bad[k] = {"count": n, "first_flat_idx": idx}
        if bad:
            report["mismatching_reps"] += 1
            ev = {"rep": r, "diff": bad}
            if "point_list" in bad:
                i0 = bad["point_list"]["first_flat_idx"][0]
                rng = first["ranges"]
                tile = int(((rng[:, 0] <= i0) & (rng[:, 1] > i0)).nonzero().reshape(-1)[0])
                lo, hi = int(rng[tile, 0]), int(rng[tile, 1])
                ev["tile"] = tile
                ev["tile_len"] = hi - lo
                sl = slice(max(lo, i0 - 3), min(hi, i0 + 5))
                ids_a = first["point_list"][sl].long()
                ids_b = cur["point_list"][sl].long()
                ev["ids_first"] = ids_a.tolist()
                ev["ids_now"] = ids_b.tolist()
                ev["depth_bits_first"] = first["rec"][ids_a, 6].view(torch.int32).tolist()
                ev["depth_bits_now"] = first["rec"][ids_b, 6].view(torch.int32).tolist()
            elif "color" in bad:
                i0 = bad["color"]["first_flat_idx"][0] % (sc.W * sc.H)
                y, x = divmod(i0, W)
                ev["pixel"] = [x, y]
                ev["tile"] = (y // 16) * gx + x // 16
                ev["values"] = [float(first["color"].reshape(3, -1)[0, i0]), float(cur["color"].reshape(3, -1)[0, i0])]
            if len(report["events"]) < 6:
                report["events"].append(ev)
    print("DET_RESULT " + json.dumps(report), flush=True)


def main():
    out = {}
    for workload, procs, reps in (("C3", 4, 150), ("C2", 2, 150)):
        for p in range(procs):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", workload, str(reps)], stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True, timeout=600)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("DET_RESULT ")]
            res = json.loads(line[0][len("DET_RESULT "):]) if line else {"error": r.stdout[-1500:]}
            out[f"{workload}#{p}"] = res
            ref = out[f"{workload}#0"].get("sha_first")
            if ref and res.get("sha_first"):
                res["differs_from_process_0"] = [k for k in ref if ref[k] != res["sha_first"].get(k)]
            print(workload, p, json.dumps(res)[:3000], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "determinism.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if "--worker" in sys.argv:
        i = sys.argv.index("--worker")
        worker(sys.argv[i + 1], int(sys.argv[i + 2]))
    else:
        main()
