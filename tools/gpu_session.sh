#!/bin/bash
# Development helper: what one `gpurun -- bash tools/gpu_session.sh` call runs.  EVERY step has its own short timeout: a kernel
# that hangs must cost a minute, not the round's GPU budget (it did once: profiles/r01_SUMMARY.md, "two lanes per G2 bucket").
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 180 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/s_tests.log
[ $rc -ne 0 ] && exit $rc        # nothing else is queued behind a failing or hanging test run
timeout 240 python bench.py --no-cpu > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/s_bench.json
