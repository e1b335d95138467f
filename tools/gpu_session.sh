#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_groth16.py -x -q -s > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/s_tests.log
timeout 900 python bench.py --no-cpu --scalars witness > gpurun_out/s_bench_witness.json 2> gpurun_out/s_bench_w.err; echo "bench witness rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s_bench_witness.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_breakdown'], d['proof_sha'])
PY
tail -2 gpurun_out/s_bench_w.err
