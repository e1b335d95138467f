#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/s_tests.log
