#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k cfg5 2>&1 | grep -E "fullsize|passed|failed" | cut -c1-600; done
timeout 300 python tools/bench_binning.py cfg2 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','gpu_launches')}, 'e2e', d['e2e']['value'])
print({k:round(v,4) for k,v in d['kernels_ms'].items()})
PY
