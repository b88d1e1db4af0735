#!/bin/bash
# ncu --set full captures of the secondary workloads' hot kernels: cfg5 (k_itemtile<PREDICT|FILTER>, k_topk of one recommend
# call) and cfg3 (one title step).  Reports + raw CSVs come back in gpurun_out/ (64 MiB return budget: ~3 MB per launch).
mkdir -p gpurun_out
C5="python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_itemtile|k_topk|k_thr' -s ${C5_SKIP:-7} -c ${C5_COUNT:-7} -f -o gpurun_out/prof_cfg5 $C5 > gpurun_out/ncu_cfg5_full.log 2>&1; echo "cfg5 rc=$?"
ncu -i gpurun_out/prof_cfg5.ncu-rep --page raw --csv > gpurun_out/prof_cfg5.raw.csv 2>/dev/null
if [ -z "$SKIP_CFG3" ]; then
C3="python bench.py --workload cfg3 --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_title|k_dw_adam|k_charcnn|k_dh|k_mix' -s ${C3_SKIP:-12} -c ${C3_COUNT:-10} -f -o gpurun_out/prof_cfg3 $C3 > gpurun_out/ncu_cfg3_full.log 2>&1; echo "cfg3 rc=$?"
ncu -i gpurun_out/prof_cfg3.ncu-rep --page raw --csv > gpurun_out/prof_cfg3.raw.csv 2>/dev/null
fi
[ "$(du -sm gpurun_out | cut -f1)" -gt 58 ] && rm -f gpurun_out/prof_cfg3.ncu-rep
ls -la gpurun_out | grep prof_cfg
