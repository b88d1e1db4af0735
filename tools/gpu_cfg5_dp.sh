#!/bin/bash
# item-sharded challenge inference on N GPUs (cfg5)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 2 > gpurun_out/bench_cfg5_dp$N.json 2> gpurun_out/bench_cfg5_dp$N.err; echo "rc=$?"
tail -c 900 gpurun_out/bench_cfg5_dp$N.json; tail -5 gpurun_out/bench_cfg5_dp$N.err | cut -c1-300
