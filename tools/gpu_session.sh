#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s_tests.log
timeout 600 python tools/microbench.py g2 > gpurun_out/s_mb_g2.log 2>&1; echo "mb rc=$?"; grep -E "msm_g" gpurun_out/s_mb_g2.log
timeout 900 python bench.py --no-cpu > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_breakdown'])
PY
