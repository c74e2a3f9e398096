"""Where the eager end-to-end loop of bench.py loses time on the device: the same double-buffered loop (run_e2e) under
torch.profiler, then the idle gaps between consecutive operations of the compute stream, aggregated by (previous op -> next op).

    python tools/e2e_gaps.py            # writes gpurun_out/e2e_gaps.json
"""
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "4dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    os.environ["G4R_E2E_NO_GRAPH"] = "1"
    import bench
    import diff_gaussian_rasterization as dgr
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = bench.Workload("C3", 0, dev)
    bench.run_e2e(dgr, wl, 20, 5, lambda: None)                      # warm: allocator, capacity hint
    steps = 30
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        bench.run_e2e(dgr, wl, steps, 3, lambda: None)
    path = os.path.join(ROOT, "gpurun_out", "e2e_trace.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    os.remove(path)
    by_stream = defaultdict(list)
    for e in ev:
        by_stream[e["args"].get("stream")].append(e)
    main_stream = max(by_stream, key=lambda s: sum(1 for e in by_stream[s] if e["cat"] == "kernel"))
    ops = sorted(by_stream[main_stream], key=lambda e: e["ts"])
    # steady state: from the first composite_backward of the timed loop's 5th step on
    names = [e["name"] for e in ops]
    starts = [i for i, n in enumerate(names) if "project_kernel" in n]
    ops = ops[starts[8]:starts[-2]]
    n_steps = sum(1 for e in ops if "project_kernel" in e["name"])
    span = ops[-1]["ts"] + ops[-1]["dur"] - ops[0]["ts"]
    busy = sum(e["dur"] for e in ops)
    gaps = defaultdict(lambda: [0.0, 0])
    short = lambda n: n.replace("void ", "").split("(")[0].split("<")[0][-60:]
    for a, b in zip(ops, ops[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        if g > 1.0:
            k = f"{short(a['name'])} -> {short(b['name'])}"
            gaps[k][0] += g
            gaps[k][1] += 1
    top = sorted(gaps.items(), key=lambda kv: -kv[1][0])[:25]
    copies = [e for s, l in by_stream.items() if s != main_stream for e in l if e["cat"] == "gpu_memcpy" and e["dur"] > 100]
    rep = dict(steps=n_steps, span_us_per_step=span / n_steps, busy_us_per_step=busy / n_steps, idle_us_per_step=(span - busy) / n_steps,
               h2d_copy_us=[round(e["dur"], 1) for e in copies[:8]],
               top_gaps_us_per_step={k: dict(us_per_step=round(v[0] / n_steps, 2), count_per_step=round(v[1] / n_steps, 2)) for k, v in top},
               ops_per_step=len(ops) / n_steps, host=dgr.host_backend())
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "e2e_gaps.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1)[:6000])


if __name__ == "__main__":
    main()
