#!/bin/bash
# Round-2 second profile pass (one GPU): ncu --set full of the tiled kernels in their own configurations.
# Every ncu run is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pytest.log
C5="python scripts/bench_configs.py --only c5"
timeout 300 $C5 > gpurun_out/plain_c5_r02.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jacobi3d -s 6 -c 2 -f -o gpurun_out/prof_c5_r02 $C5 > gpurun_out/ncu_c5_r02.log 2>&1
echo "c5 rc=$?"; cat gpurun_out/plain_c5_r02.log
S="python scripts/load_probe_raw.py 1373"
timeout 300 $S > gpurun_out/plain_shard_r02.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ratio4|ingest_bands|fill_leaf' -s 4 -c 8 -f -o gpurun_out/prof_shard_r02 $S > gpurun_out/ncu_shard_r02.log 2>&1
echo "shard rc=$?"; tail -1 gpurun_out/plain_shard_r02.log
ls -la gpurun_out/*.ncu-rep
