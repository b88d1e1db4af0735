#!/bin/bash
# ncu evidence: launch list of a short bench run + full captures of the dominant kernels.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adam_vec4 -s 4 -c 2 -o gpurun_out/prof_adam $B > gpurun_out/ncu_adam.log 2>&1; echo "adam rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_itemtile -s 6 -c 2 -o gpurun_out/prof_itemtile $B > gpurun_out/ncu_itemtile.log 2>&1; echo "itemtile rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dh -s 3 -c 1 -o gpurun_out/prof_dh $B > gpurun_out/ncu_dh.log 2>&1; echo "dh rc=$?"
ls -la gpurun_out/
