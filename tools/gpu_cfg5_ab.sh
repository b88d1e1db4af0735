#!/bin/bash
# cfg5 A/B on ONE box: one bench run per entry of FLAGS (dae_model_set_debug bits), then optionally an ncu --set full
# capture of the full-range FILTER launch of the first variant.  usage: FLAGS="0 65536" [NCU=1] tools/gpu_cfg5_ab.sh
mkdir -p gpurun_out
for f in ${FLAGS:-0}; do
  timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 2 --debug-flags $f > gpurun_out/bench_cfg5_f$f.json 2> gpurun_out/bench_cfg5_f$f.err || tail -3 gpurun_out/bench_cfg5_f$f.err
  python - $f <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_cfg5_f%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("flags", sys.argv[1], "ms/call %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phase_ms"].items()})
PY
done
if [ -n "$NCU" ]; then
  C5="python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline --debug-flags ${NCU_FLAGS:-0}"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_itemtile' -s ${C5_SKIP:-5} -c 1 -f -o gpurun_out/prof_filter $C5 > gpurun_out/ncu_filter.log 2>&1; echo "ncu rc=$?"
  ncu -i gpurun_out/prof_filter.ncu-rep --page raw --csv > gpurun_out/prof_filter.raw.csv 2>/dev/null
  ls -la gpurun_out | grep prof_filter
fi
