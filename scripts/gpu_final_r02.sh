#!/bin/bash
# Round-2 closing run on one GPU: tests, smoke, bench (BENCH-style line), reference arm, other configs, ncu launch list.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -2 gpurun_out/r2_final_bench_n1.err | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
python scripts/bench_configs.py > gpurun_out/r2_final_configs.jsonl 2> gpurun_out/r2_final_configs.err; cat gpurun_out/r2_final_configs.jsonl
B="python bench.py --steps 2 --warmup 1 --e2e-steps 2 --roofline-launches 8 --no-cpu-baseline --no-scene --no-variants"
timeout 300 $B > gpurun_out/plain_r02.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_list_r02.log 2>&1
echo "list rc=$?"
S="python scripts/load_probe_raw.py 1373"
timeout -k 10 300 $S > gpurun_out/plain_shard_r02.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'ratio_bulk|ingest_bands|fill_leaf' -s 4 -c 6 -f -o gpurun_out/prof_shard_r02 $S > gpurun_out/ncu_shard_r02.log 2>&1
echo "shard rc=$?"; tail -1 gpurun_out/plain_shard_r02.log
