#!/bin/bash
mkdir -p gpurun_out
echo "--- new"; timeout 300 python tools/bench_binning.py cfg2 --kernels --save /tmp/new2.pt 2>&1 | tail -16
echo "--- classic"; GS3D_SORT=classic timeout 300 python tools/bench_binning.py cfg2 --kernels --save /tmp/old2.pt 2>&1 | tail -22
echo "--- cfg5 new"; timeout 300 python tools/bench_binning.py cfg5 --save /tmp/new5.pt 2>&1 | tail -2
echo "--- cfg5 classic"; GS3D_SORT=classic timeout 300 python tools/bench_binning.py cfg5 --save /tmp/old5.pt 2>&1 | tail -2
python - <<'PY'
import torch
for t in ('2','5'):
    a=torch.load(f'/tmp/new{t}.pt'); b=torch.load(f'/tmp/old{t}.pt')
    print('cfg'+t, {k: bool(torch.equal(a[k], b[k])) for k in a}, 'ids differing', int((a['ids']!=b['ids']).sum()))
PY
for i in 1 2; do GS3D_SORT=classic timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k cfg5 2>&1 | tail -2; done
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k cfg5 2>&1 | tail -2; done
