#!/bin/bash
# ncu evidence: launch list of a short bench run + full captures of the dominant kernels.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
for K in k_adam_rows_vec4 k_itemtile k_dw_adam k_dh k_encode_fwd k_scatter_shard; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_$K $B > gpurun_out/ncu_$K.log 2>&1; echo "$K rc=$?"
  ncu -i gpurun_out/prof_$K.ncu-rep --page raw --csv > gpurun_out/prof_$K.raw.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/ | grep -E "prof_|launches"
