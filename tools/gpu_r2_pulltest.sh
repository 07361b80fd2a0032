#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_parallel.py -q -s -k "sparse_gradient_exchange" 2>&1 | grep -E "dp rank|passed|failed" ; done
echo "--- classic sort"
GS3D_SORT=classic timeout 600 python -m pytest tests/test_gpu_parallel.py -q -s -k "sparse_gradient_exchange" 2>&1 | grep -E "dp rank|passed|failed"
