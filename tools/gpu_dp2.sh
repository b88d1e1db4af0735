#!/bin/bash
# 2-GPU call: IPC / NVLink probe, then the data-parallel bench.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== p2p probe"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_probe.py > gpurun_out/p2p_probe.log 2>&1; echo "rc=$?"; grep -E "rank|rror" gpurun_out/p2p_probe.log | tail -8
echo "== bench dp2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_dp2.json; tail -5 gpurun_out/bench_dp2.err
