#!/bin/bash
# quick GPU visit: parity tests + our bench arm (+ optional extra command in $1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','kernels_ms')}, d['e2e']['value'])
PY
if [ -n "$1" ]; then bash -c "$1"; fi
