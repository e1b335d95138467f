#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py tests/test_gpu_verify.py -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s_tests.log
for v in 0 4 3 2; do
ZKPOR_G2_PAIR=$v timeout 600 python tools/microbench.py g2 > gpurun_out/s_mb_pair$v.log 2>&1; echo "pair=$v rc=$?"; grep -E "msm_g2" gpurun_out/s_mb_pair$v.log
done
