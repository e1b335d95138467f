#!/bin/bash
# development helper: one gpurun call = the whole GPU test suite, the contract benchmark and its ncu launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s_tests.log
timeout 900 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; cat gpurun_out/s_bench.json | cut -c1-1500; tail -3 gpurun_out/s_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_logn22.csv python bench.py --steps 2 --warmup 1 --log-n 22 --no-e2e --no-cpu > gpurun_out/s_bench_ncu.log 2>&1; echo "ncu rc=$?"
