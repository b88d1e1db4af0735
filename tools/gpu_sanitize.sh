#!/bin/bash
# compute-sanitizer memcheck over one small parity test (argument: pytest -k expression)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dp.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "$1" > gpurun_out/sanitize.log 2>&1; echo "rc=$?"
grep -vE "^\s*$" gpurun_out/sanitize.log | grep -E "=========|Error|error|passed|failed" | head -60
