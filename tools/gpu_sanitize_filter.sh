#!/bin/bash
# compute-sanitizer memcheck, then racecheck, over the smallest fused decode + top-K parity test (the FILTER epilogue's
# shared-memory queue / staging halves, the list appends); logs to gpurun_out/sanitize_filter_{memcheck,racecheck}.log
mkdir -p gpurun_out
T='tests/test_gpu_model.py::test_fused_decode_topk_equals_dense_ranking[9000-8000-64-64-100]'
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-100} compute-sanitizer --tool $tool --print-limit 12 python -m pytest "$T" -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitize_filter_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "=========|passed|failed" gpurun_out/sanitize_filter_$tool.log | grep -vE "^=========\s*$" | cut -c1-220 | head -14
done
