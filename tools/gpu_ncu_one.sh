#!/bin/bash
# ncu --set full capture of one kernel (regex $1, skip $2 launches) of a short bench run ($NCU_CMD overrides the
# command); raw / details CSV and the report come back in gpurun_out/.
mkdir -p gpurun_out
B=${NCU_CMD:-"python bench.py --steps 3 --warmup 3 --no-cpu-baseline"}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-4} -c 1 -f -o gpurun_out/prof_$1 $B > gpurun_out/ncu_$1.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$1.ncu-rep --page details --csv > gpurun_out/prof_$1.details.csv 2>/dev/null
ls -la gpurun_out | grep prof_$1
