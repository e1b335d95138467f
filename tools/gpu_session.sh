#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_groth16.py tests/test_gpu_verify.py -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s_tests.log
for ov in 0 1; do
ZKPOR_OVERLAP_NTT=$ov timeout 900 python bench.py --no-cpu > gpurun_out/s_bench_ov$ov.json 2> gpurun_out/s_bench_ov$ov.err; echo "bench overlap=$ov rc=$?"; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/s_bench_ov$ov.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['kernel_breakdown'], d['roofline']['launch_ms'], d['proof_sha'])
PY
done
