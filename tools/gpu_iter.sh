#!/bin/bash
# One optimisation iteration on the GPU: parity tests, then a bench run.  Logs to gpurun_out/.
mkdir -p gpurun_out
echo "== pytest kernels"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_kernels.log
echo "== pytest model"; timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_model.log
echo "== pytest dp"; timeout 1200 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_dp.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 10 ${BENCH_ARGS} > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
    print("value %.0f playlists/s  ms/step %.4f  e2e %.0f (%.4f ms)  launches %d" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
    print("roofline", d['roofline'])
    print("phases", {k: round(v, 4) for k, v in d['phase_ms'].items()})
    print("clocks", d['clocks'], "cpu", d.get('cpu_baseline', {}).get('value'))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench_iter.err
