"""Summarise an `ncu --set full` report (brought back in gpurun_out/) into profiles/: a markdown table per kernel and a
small JSON with DRAM bytes per launch that bench.py reports as `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep r01_v5
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads/instr"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full summary ({tag})", "",
           f"Source: `{os.path.basename(rep)}` captured under gpurun with `--clock-control none --import-source on` while running "
           "`python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (workload C3: 500k Gaussians, 640x480). Per-launch values; ncu "
           "serialises kernels and flushes caches between replays, so compare SHARES with the live numbers of bench.py, not absolutes.", ""]
    traffic = {}
    names = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        names.append(name)
        out.append(f"## {name}")
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        for m, label in METRICS:
            if m in idx:
                out.append(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic[name] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr}
        if "smsp__inst_executed.sum" in idx:
            traffic[name]["warp_instructions"] = float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))
        out.append(f"| DRAM traffic per launch | {(rd + wr) / 1e6:.2f} MB |")
        out.append("")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump({"tag": tag, "workload": "C3", "kernels": traffic}, f, indent=1)
    print("wrote", len(names), "kernels:", names)


if __name__ == "__main__":
    main()
