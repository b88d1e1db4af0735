#!/bin/bash
# One full GPU pass: every -m gpu test (as the driver runs them), smoke(), the default bench line, then the ncu
# evidence for profiles/: launch list of a short bench run + ONE --set full run capturing one launch of each hot kernel.
# Everything logs to gpurun_out/.  SKIP_TESTS=1 / SKIP_NCU=1 shorten the call; SOAK=1 adds an 8000-step stability run.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest ${PYTEST_ARGS:-tests/} -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
  echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
fi
echo "== bench"; timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench_round.json 2> gpurun_out/bench_round.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_round.json').read().strip().splitlines()[-1])
    print("value %.0f playlists/s  ms/step %.4f  e2e %.0f (%.4f ms)  launches %d" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
    print("roofline", d['roofline'])
    print("phases", {k: round(v, 4) for k, v in d['phase_ms'].items()})
    print("clocks", d['clocks'], "cpu", d.get('cpu_baseline', {}).get('value'))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench_round.err
if [ -n "$SOAK" ]; then
  # long runs catch what parity tests cannot: the round-1 staging bug of k_dw_adam_fused showed once in several thousand steps
  echo "== soak"; timeout 600 python bench.py --steps ${SOAK_STEPS:-8000} --warmup 20 --no-cpu-baseline > gpurun_out/bench_soak.json 2> gpurun_out/bench_soak.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_soak.json; grep -o "device trap[^]]*" gpurun_out/bench_soak.err | head -3
fi
if [ -z "$SKIP_NCU" ]; then
  B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
  # hot kernels of one step, in launch order; skip the first 4 steps' worth of matches, capture one step's worth (+ slack)
  KRE='k_adam_rows_vec4|k_adam_listed|k_itemtile|k_dw_adam|k_dh|k_encode_fwd|k_scatter_shard|k_scatter_det|k_da_own'
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s ${NCU_SKIP:-36} -c ${NCU_COUNT:-9} -f -o gpurun_out/prof_hot $B > gpurun_out/ncu_hot.log 2>&1; echo "full capture rc=$?"
  ncu -i gpurun_out/prof_hot.ncu-rep --page raw --csv > gpurun_out/prof_hot.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_hot.ncu-rep --page details --csv > gpurun_out/prof_hot.details.csv 2>/dev/null
  # the report itself only travels back when it fits the 64 MiB return budget
  [ "$(stat -c%s gpurun_out/prof_hot.ncu-rep 2>/dev/null || echo 0)" -gt 40000000 ] && rm -f gpurun_out/prof_hot.ncu-rep
  ls -la gpurun_out/ | grep -E "prof_|launches"
fi
