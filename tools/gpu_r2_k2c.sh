#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py::test_cfg5_4k_forward > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
echo "--- new"; timeout 300 python tools/bench_binning.py cfg2 --kernels --save /tmp/new2.pt 2>&1 | grep -v Warn | tail -18
echo "--- classic"; GS3D_SORT=classic timeout 300 python tools/bench_binning.py cfg2 --save /tmp/old2.pt 2>&1 | tail -2
echo "--- cfg5 new"; timeout 300 python tools/bench_binning.py cfg5 --save /tmp/new5.pt 2>&1 | tail -2
echo "--- cfg5 classic"; GS3D_SORT=classic timeout 300 python tools/bench_binning.py cfg5 --save /tmp/old5.pt 2>&1 | tail -2
python - <<'PY'
import torch
for t in ('2','5'):
    a=torch.load(f'/tmp/new{t}.pt'); b=torch.load(f'/tmp/old{t}.pt')
    print('cfg'+t, {k: bool(torch.equal(a[k], b[k])) for k in a}, 'ids differing', int((a['ids']!=b['ids']).sum()))
PY
