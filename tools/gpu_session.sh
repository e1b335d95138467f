#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s_tests.log
timeout 600 python tools/microbench.py msm g2 check > gpurun_out/s_mb.log 2>&1; echo "mb rc=$?"; grep -E "msm_g|check" gpurun_out/s_mb.log
timeout 900 python bench.py --no-cpu > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_breakdown'], d['proof_sha'])
PY
