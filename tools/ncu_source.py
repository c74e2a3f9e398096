"""Summarise the SASS-level source page of an ncu report for one kernel (run here, no GPU needed):

    python tools/ncu_source.py gpurun_out/prof_all.ncu-rep composite_backward [--top 40] [--md out.md]

Prints the warp-stall totals by reason, the instruction mix (opcode -> warp instructions executed) and the hottest SASS
instructions by stall samples.  Used for the per-kernel stall tables committed under profiles/."""
from __future__ import annotations

import csv
import io
import subprocess
import sys
from collections import Counter


def load(rep: str, kernel: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    blocks, cur, name = [], None, None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            name = next(csv.reader([line]))[1]
            cur = []
            blocks.append((name, cur))
        elif cur is not None:
            cur.append(line)
    res = []
    for name, lines in blocks:
        rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
        res.append((name, rows))
    return res


def num(v):
    try:
        return float(v)
    except (TypeError, ValueError):
        return 0.0


def summarise(name, rows, top=40):
    lines = [f"## {name}", ""]
    tot_inst = sum(num(r["Instructions Executed"]) for r in rows)
    tot_samp = sum(num(r["# Samples"]) for r in rows)
    lines.append(f"SASS instructions: {len(rows)}; warp instructions executed: {tot_inst:.0f}; stall samples: {tot_samp:.0f}")
    reasons = [k for k in rows[0].keys() if k.startswith("stall_") and "(Not Issued)" not in k]
    rs = Counter({k: sum(num(r[k]) for r in rows) for k in reasons})
    lines += ["", "| stall reason | samples | share |", "|---|---|---|"]
    for k, v in rs.most_common():
        if v > 0:
            lines.append(f"| {k} | {v:.0f} | {100 * v / max(tot_samp, 1):.1f} % |")
    mix = Counter()
    for r in rows:
        op = r["Source"].strip().split()
        if not op:
            continue
        o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        mix[o.split(".")[0]] += num(r["Instructions Executed"])
    lines += ["", "| opcode | warp instructions | share |", "|---|---|---|"]
    for k, v in mix.most_common(24):
        lines.append(f"| {k} | {v:.0f} | {100 * v / max(tot_inst, 1):.1f} % |")
    lines += ["", f"Hottest {top} instructions by stall samples:", "", "| # | samples | executed | top stall | SASS |", "|---|---|---|---|---|"]
    order = sorted(range(len(rows)), key=lambda i: -num(rows[i]["# Samples"]))[:top]
    for i in sorted(order):
        r = rows[i]
        best = max(reasons, key=lambda k: num(r[k]))
        lines.append(f"| {i} | {num(r['# Samples']):.0f} | {num(r['Instructions Executed']):.0f} | {best[6:]} | `{r['Source'].strip()}` |")
    return "\n".join(lines) + "\n"


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    text = ""
    for name, rows in load(rep, kernel):
        text += summarise(name, rows, top) + "\n"
    if "--md" in sys.argv:
        with open(sys.argv[sys.argv.index("--md") + 1], "w") as f:
            f.write(text)
    print(text)
    if "--dump" in sys.argv:
        for name, rows in load(rep, kernel):
            for i, r in enumerate(rows):
                print(f"{i:5d} {num(r['# Samples']):7.0f} {num(r['Instructions Executed']):10.0f}  {r['Source'].strip()}")


if __name__ == "__main__":
    main()
