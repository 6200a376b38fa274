#!/bin/bash
# Round-2 profiles (one GPU).  Every ncu run is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --e2e-steps 2 --roofline-launches 8 --no-cpu-baseline --no-scene --no-variants"
timeout 300 $B > gpurun_out/plain_r02.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_list_r02.log 2>&1
echo "list rc=$?"
R="python scripts/res_full_probe.py 512 2"
timeout 200 $R > gpurun_out/plain_res_r02.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resident2d -s 3 -c 3 -f -o gpurun_out/prof_res_r02 $R > gpurun_out/ncu_res_r02.log 2>&1
echo "resident rc=$?"
tail -3 gpurun_out/plain_res_r02.log
ls -la gpurun_out/prof_res_r02.ncu-rep
