#!/bin/bash
# Multi-GPU pass (gpurun --gpus N -- 'bash tools/gpu_dp_check.sh N'): the one-process-per-GPU IPC parity test, the N-rank
# train bench (its self-check against one rank precedes the timing) and the item-sharded challenge inference bench.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_dp.py -x -q -m gpu -p no:cacheprovider -k "ipc" 2>&1 | tail -4
fi
P=$((29500 + RANDOM % 200))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps ${STEPS:-200} --warmup 20 $BENCH_ARGS > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "train rc=$?"
tail -c 1500 gpurun_out/bench_dp$N.json; grep -o "device trap[^]]*" gpurun_out/bench_dp$N.err | head -3; tail -3 gpurun_out/bench_dp$N.err
if [ -z "$SKIP_CFG5" ]; then
  P=$((29700 + RANDOM % 200))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_dp$N.json 2> gpurun_out/bench_cfg5_dp$N.err; echo "cfg5 rc=$?"
  tail -c 900 gpurun_out/bench_cfg5_dp$N.json; tail -3 gpurun_out/bench_cfg5_dp$N.err
fi
