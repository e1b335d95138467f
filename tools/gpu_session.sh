#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_logn22_v4.csv python bench.py --steps 2 --warmup 1 --log-n 22 --no-e2e --no-cpu > gpurun_out/s_bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_accumulate<|k_part_sort|k_view_lists" -s 2 -c 6 -o gpurun_out/r01_v4_kernels -f python bench.py --steps 1 --warmup 0 --log-n 24 --no-e2e --no-cpu > gpurun_out/s_bench_ncu2.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s_bench_ref.json 2> gpurun_out/s_bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/s_bench_ref.json
