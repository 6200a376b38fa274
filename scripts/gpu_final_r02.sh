#!/bin/bash
# Round-2 closing run on one GPU: tests, smoke, bench (BENCH-style line), other configs, ncu of the resident kernel's classes.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -2 gpurun_out/r2_final_bench_n1.err | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
python scripts/bench_configs.py > gpurun_out/r2_final_configs.jsonl 2> gpurun_out/r2_final_configs.err; cat gpurun_out/r2_final_configs.jsonl
R="python scripts/res_full_probe.py 512 2"
timeout 200 $R > gpurun_out/plain_res_r02b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resident2d -s 4 -c 4 -f -o gpurun_out/prof_res_r02b $R > gpurun_out/ncu_res_r02b.log 2>&1
echo "resident ncu rc=$?"; tail -2 gpurun_out/plain_res_r02b.log
