#!/bin/bash
# GPU parity tests (kernel level, then model level) in separate processes; logs to gpurun_out/.
mkdir -p gpurun_out
echo "== pytest kernels"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_kernels.log
echo "== pytest model"; timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/pytest_model.log
