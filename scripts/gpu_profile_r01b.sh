#!/bin/bash
# Profiles of the final round-1 kernels (one GPU): launch list of the default bench path on a reduced batch and a
# full capture of the resident kernel.  Every ncu run is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --patches 64 --steps 1 --warmup 0 --e2e-steps 1 --roofline-launches 4 --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_r01b.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01b.csv $B > gpurun_out/ncu_list_r01b.log 2>&1
echo "list rc=$?"
R="python scripts/res_probe.py 16"
timeout 100 $R > gpurun_out/plain_res_r01b.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resident2d -s 1 -c 1 -f -o gpurun_out/prof_res_r01b $R > gpurun_out/ncu_res_r01b.log 2>&1
echo "resident rc=$?"
tail -2 gpurun_out/plain_res_r01b.log
