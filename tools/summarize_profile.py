#!/usr/bin/env python
"""Turn the ncu CSVs of tools/gpu_round.sh into the tracked evidence under profiles/.

    python tools/summarize_profile.py --tag r01b [--src gpurun_out] [--note "..."]

Writes profiles/<tag>_launches.csv (launch list aggregated by kernel), profiles/<tag>_hot_raw.csv (the raw page of
the --set full capture, key metrics only), profiles/<tag>_summary.md, and profiles/traffic.json (DRAM bytes per
launch of every captured kernel; bench.py quotes the dominant kernel's figure as roofline.traffic)."""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
       "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--src", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--note", default="")
    ap.add_argument("--bench", default="bench_round.json")
    a = ap.parse_args()
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)

    # ---- launch list ----
    agg = collections.OrderedDict()
    with open(os.path.join(a.src, "launches.csv")) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ns)
    total = sum(t for _, t in agg.values())
    ranked = sorted(agg.items(), key=lambda kv: -kv[1][1])
    with open(os.path.join(out, a.tag + "_launches.csv"), "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, t) in ranked:
            f.write("%s,%d,%.1f,%.2f,%.4f\n" % (k, n, t / 1e3, t / 1e3 / n, t / total))

    # ---- full capture ----
    rows = list(csv.reader(open(os.path.join(a.src, "prof_hot.raw.csv"))))
    rows = [r for r in rows if len(r) > 20]
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [c for c in KEY if c in hdr]
    traffic = {}
    table = []
    for r in data:
        k = short(r[ki])
        rec = {"kernel": k}
        for c in cols:
            i = hdr.index(c)
            rec[c] = (r[i], units[i])
        rd = to_bytes(*rec["dram__bytes_read.sum"])
        wr = to_bytes(*rec["dram__bytes_write.sum"])
        rec["dram_bytes"] = rd + wr
        table.append(rec)
        traffic.setdefault(k, []).append({"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
                                          "gpu_time_us": float(rec["gpu__time_duration.sum"][0].replace(",", ""))
                                          * (1e-3 if rec["gpu__time_duration.sum"][1] in ("ns", "nsecond") else 1.0)})
    with open(os.path.join(out, a.tag + "_hot_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + cols)
        w.writerow([""] + [units[hdr.index(c)] for c in cols])
        for rec in table:
            w.writerow([rec["kernel"]] + [rec[c][0] for c in cols])
    with open(os.path.join(out, "traffic.json"), "w") as f:
        json.dump({"source": "profiles/%s_hot_raw.csv (ncu --set full --clock-control none, one launch each, "
                             "python bench.py --steps 3 --warmup 3 --no-cpu-baseline)" % a.tag, "kernels": traffic}, f, indent=1)

    bench = None
    try:
        bench = json.loads(open(os.path.join(a.src, a.bench)).read().strip().splitlines()[-1])
    except Exception:
        pass

    with open(os.path.join(out, a.tag + "_summary.md"), "w") as f:
        f.write("# ncu evidence `%s` (B200, `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`)\n\n" % a.tag)
        if a.note:
            f.write(a.note + "\n\n")
        f.write("Captured with `tools/gpu_round.sh`: (1) `ncu --metrics gpu__time_duration.sum --clock-control none` launch list, "
                "(2) one `ncu --set full --clock-control none --import-source on` run capturing one launch of every hot kernel of a step.\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("## Launch list (all launches of the run, aggregated by kernel)\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in ranked:
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, n, t / 1e3, t / 1e3 / n, 100 * t / total))
        if bench:
            ph = bench.get("phase_ms", {})
            f.write("\nbench.py of the same build (not under ncu): %.4f ms/step, %.0f playlists/s device-timed, e2e %.0f playlists/s; "
                    "CUDA-event phases (ms): %s\n" % (bench["ms_per_step"], bench["value"], bench["e2e"]["value"],
                                                    ", ".join("%s %.3f" % (k, v) for k, v in ph.items())))
        f.write("\n## Full captures (`--set full`), key raw metrics, one launch each\n\n")
        f.write("| kernel | " + " | ".join(c.split(".")[0] for c in cols) + " | DRAM MB |\n|" + "---|" * (len(cols) + 2) + "\n")
        for rec in table:
            f.write("| `%s` | " % rec["kernel"] + " | ".join("%s %s" % rec[c] for c in cols) + " | %.1f |\n" % (rec["dram_bytes"] / 1e6))
    print("wrote profiles/%s_{launches.csv,hot_raw.csv,summary.md} and profiles/traffic.json" % a.tag)


if __name__ == "__main__":
    main()
