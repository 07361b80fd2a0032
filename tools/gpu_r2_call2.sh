#!/bin/bash
# round 2, GPU visit 2: whole-path parity after the rounding-exact projection, route-2 test, full GPU suite
mkdir -p gpurun_out
timeout 900 python tools/parity_fullsize.py cfg3 bg posed cfg2 cfg5 > gpurun_out/parity_fullsize.jsonl 2> gpurun_out/parity_fullsize.err; echo "parity rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/parity_fullsize.jsonl'):
    d=json.loads(l)
    print({k:d[k] for k in ('workload','n_dub_ours','n_dub_ref','mask_mismatch','bits_svec','bits_alpha','bits_mean2d','bits_cov','bits_depth','rect_mismatch','ranges_equal','keys_equal','ids_tie_only','image_max_abs','image_gt_1e4','pass')})
PY
tail -3 gpurun_out/parity_fullsize.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','kernels_ms')}, d['e2e']['value'])
PY
