#!/usr/bin/env python
"""ncu --set full captures of the secondary workloads (tools/gpu_ncu_aux.sh: gpurun_out/prof_cfg5.raw.csv, prof_cfg3.raw.csv)
-> profiles/<tag>_cfg5_cfg3.md: one row per captured launch with the key raw metrics.

    python tools/summarize_aux_profile.py --tag r02c"""
import argparse
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
       ("launch__registers_per_thread", "regs"),
       ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
       ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
       ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
       ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
       ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
       ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("smsp__inst_executed.sum", "warp inst")]


def table(path):
    rows = list(csv.DictReader(open(path)))
    units, rows = rows[0], rows[1:]
    out = ["| kernel | " + " | ".join(n for _, n in KEY) + " |", "|---|" + "---|" * len(KEY)]
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dae::", "").strip()
        cells = []
        for k, _ in KEY:
            v = r.get(k, "")
            try:
                f = float(v.replace(",", ""))
                v = ("%.4g" % f) + (" " + units.get(k, "") if units.get(k, "") not in ("", "%") else "")
            except ValueError:
                pass
            cells.append(v)
        out.append("| `" + name + "` | " + " | ".join(cells) + " |")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--src", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    md = ["# ncu `--set full` captures of the secondary workloads (`%s`, B200, `tools/gpu_ncu_aux.sh`)" % a.tag, "",
          "One launch each, `--clock-control none`; times under ncu are cold-cache and serialised.", ""]
    for wl, cmd in (("cfg5", "python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline (one 4096-playlist x 2 M-item recommend call: "
                             "thresholds, dense prefix k_itemtile<1>, select, middle FILTER pass k_itemtile<2>, select, full-range FILTER pass, final select)"),
                    ("cfg3", "python bench.py --workload cfg3 --steps 2 --warmup 2 --no-cpu-baseline (title-mode train step)")):
        p = os.path.join(a.src, "prof_%s.raw.csv" % wl)
        if not os.path.exists(p):
            continue
        md += ["## %s: `%s`" % (wl, cmd), "", table(p), ""]
    b = os.path.join(a.src, "bench_round.json")
    if os.path.exists(b):
        d = json.loads(open(b).read().strip().splitlines()[-1])
        for wl in ("cfg5", "cfg3"):
            if wl in d:
                md += ["bench.py of the same build, `%s` (not under ncu): %.3f ms per %s, phases (ms) %s, roofline %s" % (
                    wl, d[wl]["ms_per_step"], "call" if wl == "cfg5" else "step",
                    {k: round(v, 3) for k, v in d[wl].get("phase_ms", {}).items()},
                    {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d[wl].get("roofline", {}).items() if k in ("bound", "achieved", "peak", "unit", "frac", "launch_ms")}), ""]
    out = os.path.join(ROOT, "profiles", a.tag + "_cfg5_cfg3.md")
    open(out, "w").write("\n".join(md))
    print("wrote", out)


if __name__ == "__main__":
    main()
