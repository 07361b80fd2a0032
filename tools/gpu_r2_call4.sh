#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_parity.py tests/test_train_steps.py -x -q -s > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|\[graph\]|Error" gpurun_out/pytest_graph.log | tail -12
for extra in "" "--no-graph"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $extra > gpurun_out/bench_ours$extra.json 2> gpurun_out/bench_ours$extra.err; echo "bench $extra rc=$?"; tail -3 gpurun_out/bench_ours$extra.err
python - "$extra" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/bench_ours{sys.argv[1]}.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','gpu_launches')}, 'e2e', d['e2e']['value'], d['impl_details']['step_launch'])
print({k:round(v,4) for k,v in d['kernels_ms'].items()})
PY
done
