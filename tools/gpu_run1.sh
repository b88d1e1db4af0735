#!/bin/bash
# First GPU call: descriptor probe, parity tests, smoke, short bench.  Everything logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe"; timeout 300 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -40 gpurun_out/probe.log
echo "== pytest kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_kernels.log
echo "== pytest model"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_model.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
