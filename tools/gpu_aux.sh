#!/bin/bash
mkdir -p gpurun_out
for w in cfg3 cfg5; do
  echo "== bench $w"; timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
