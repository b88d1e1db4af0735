#!/bin/bash
# cfg5 (challenge inference) bench + kernel launch list under ncu; logs to gpurun_out/
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg5 --steps ${STEPS:-5} --warmup 2 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "bench rc=$?"; tail -c 700 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --steps 1 --warmup 1 > gpurun_out/ncu_cfg5.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
lines = [l for l in open('gpurun_out/launches_cfg5.csv') if l.startswith('"')]
rows = list(csv.DictReader(lines))
seq = [(r["Kernel Name"].split("(")[0].replace("void ", "")[:40], float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"]) for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
for k, us, g in seq[-24:]:
    print("%-42s %10.1f us  grid %s" % (k, us, g))
PY
fi
