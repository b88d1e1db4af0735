#!/bin/bash
# item-sharded challenge inference on N GPUs: bench.py --workload cfg5 (its self-check against the unsharded list precedes the timing)
N=${1:-2}
mkdir -p gpurun_out
P=$((29700 + RANDOM % 200))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench_cfg5_dp$N$TAG.json 2> gpurun_out/bench_cfg5_dp$N$TAG.err; echo "cfg5 rc=$?"
tail -c 1200 gpurun_out/bench_cfg5_dp$N$TAG.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_cfg5_dp$N$TAG.err | tail -5
