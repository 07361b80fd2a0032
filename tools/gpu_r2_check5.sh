#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_binning.py cfg2 --kernels --save /tmp/new2.pt 2>&1 | grep -v -i warn | tail -20
GS3D_SORT=classic timeout 300 python tools/bench_binning.py cfg2 --save /tmp/old2.pt 2>&1 | tail -3 | head -1
timeout 300 python tools/bench_binning.py cfg5 --save /tmp/new5.pt 2>&1 | tail -3 | head -1
GS3D_SORT=classic timeout 300 python tools/bench_binning.py cfg5 --save /tmp/old5.pt 2>&1 | tail -3 | head -1
python - <<'PY'
import torch
for t in ('2','5'):
    a=torch.load(f'/tmp/new{t}.pt'); b=torch.load(f'/tmp/old{t}.pt')
    print('cfg'+t, {k: bool(torch.equal(a[k], b[k])) for k in a})
PY
