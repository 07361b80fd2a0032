#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','gpu_launches')}, 'e2e', d['e2e']['value'])
print({k:round(v,4) for k,v in d['kernels_ms'].items()})
PY
timeout 600 ncu --set full --clock-control none --nvtx --nvtx-include "profile_step/" -k regex:gs3d -o gpurun_out/prof_r2_step -f python tools/one_step.py > gpurun_out/ncu_r2_step.log 2>&1; echo "ncu step rc=$?"; tail -2 gpurun_out/ncu_r2_step.log
ls -la gpurun_out/prof_r2_step.ncu-rep
