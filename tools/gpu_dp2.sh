#!/bin/bash
# N-GPU call (N = $1, default 2): multi-process CUDA-IPC parity test, then the data-parallel bench.
N=${1:-2}
mkdir -p gpurun_out
echo "== pytest dp (ipc)"; timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dp$N.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_dp$N.log
echo "== bench dp$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-50} --warmup 10 --no-cpu-baseline > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_dp$N.json').read().strip().splitlines()[-1])
    print("value %.0f playlists/s  ms/step %.4f  e2e %.0f (%.4f ms)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
    print("phases", {k: round(v, 4) for k, v in d['phase_ms'].items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench_dp$N.err
