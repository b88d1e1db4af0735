#!/bin/bash
# N-GPU scaling evidence: cfg2 train bench at N GPUs (+ optional cfg5)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_dp$N.json').read().strip().splitlines()[-1])
    print("N=$N value %.0f playlists/s  ms/step %.4f  e2e %.0f (%.4f ms)" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
    print("phases", {k: round(v, 4) for k, v in d['phase_ms'].items()})
    print("roofline", d['roofline'])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_dp$N.err | cut -c1-300
if [ -n "$CFG5" ]; then bash tools/gpu_cfg5_dp.sh $N; fi
