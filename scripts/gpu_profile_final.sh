#!/bin/bash
# Round-end profiles (one GPU): launch list of the default bench path, full captures of the resident kernel and of
# the ratio-form tiled kernel.  Every ncu run is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --patches 64 --steps 1 --warmup 0 --e2e-steps 1 --roofline-launches 4 --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_final.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_list_final.log 2>&1
echo "list rc=$?"
R="python scripts/res_probe.py 16"
timeout 100 $R > gpurun_out/plain_res_final.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resident2d -s 1 -c 1 -f -o gpurun_out/prof_res_final $R > gpurun_out/ncu_res_final.log 2>&1
echo "resident rc=$?"
S="python bench_scene.py --size 4096 --discs 20 --steps 1 --warmup 0 --k-block 1"
timeout 100 $S > gpurun_out/plain_scene_final.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ratio_kernel -s 3 -c 2 -f -o gpurun_out/prof_ratio_final $S > gpurun_out/ncu_scene_final.log 2>&1
echo "ratio rc=$?"
tail -1 gpurun_out/plain_scene_final.log | cut -c1-200
