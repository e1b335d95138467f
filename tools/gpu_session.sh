#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py tests/test_gpu_fullsize.py -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s_tests.log
timeout 600 python tools/microbench.py msm check > gpurun_out/s_mb_part.log 2>&1; echo "mb partitioned rc=$?"; grep -E "msm_g|check" gpurun_out/s_mb_part.log
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/probe_launches.csv python tools/affine_probe.py 26 > gpurun_out/probe.log 2>&1; tail -2 gpurun_out/probe.log
