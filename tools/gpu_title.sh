#!/bin/bash
mkdir -p gpurun_out
echo "== pytest title"; timeout 1200 python -m pytest tests/test_gpu_title.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_title.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/pytest_title.log
