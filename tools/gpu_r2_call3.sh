#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bench_composite.py cfg2 10 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','kernels_ms','gpu_launches')}, d['e2e']['value'])
for k,v in d['rooflines'].items(): print(k, v['bound'], round(v['achieved'],2), v['unit'], round(v['frac'],4), v.get('pairs_contributing'))
PY
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/bench_cfg4_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_n1.json')); print({k:d[k] for k in ('value','ms_per_step','scaling','kernels_ms')}, d['e2e'], d['config'])"
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --check > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; echo "cfg5 rc=$?"; tail -3 gpurun_out/bench_cfg5_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_n1.json')); print({k:d[k] for k in ('value','ms_per_step','scaling','sharded_vs_single_gpu_max_abs')}, d['e2e'], {k:d['parity'].get(k) for k in ('pass','image_max_abs','rect_mismatch','error')})"
