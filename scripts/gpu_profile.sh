#!/bin/bash
# ncu launch list + one full capture of the stencil kernel (run under gpurun, one GPU).
# usage: gpu_profile.sh <k_block> <tag>
set -u
KB=${1:-1}; TAG=${2:-r01_k$KB}
mkdir -p gpurun_out
CMD="python bench.py --patches 64 --steps 1 --warmup 0 --e2e-steps 1 --roofline-launches 3 --no-cpu-baseline --k-block $KB"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:jacobi2d -s 6 -c 2 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -12
