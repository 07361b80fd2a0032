#!/bin/bash
# round 2, GPU visit 1: op-order probe, whole-path parity at BASELINE sizes, reference Python over the shim
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 300 python tools/probe_torch_order.py 1000000 > gpurun_out/probe_order.json 2> gpurun_out/probe_order.err; echo "probe rc=$?"
timeout 900 python tools/parity_fullsize.py cfg3 bg posed cfg2 cfg5 > gpurun_out/parity_fullsize.jsonl 2> gpurun_out/parity_fullsize.err; echo "parity rc=$?"
cat gpurun_out/parity_fullsize.jsonl | cut -c1-900
tail -3 gpurun_out/parity_fullsize.err
timeout 600 python -m pytest tests/test_gpu_reference_python.py -x -q -s > gpurun_out/pytest_refpy.log 2>&1; echo "refpy rc=$?"
tail -15 gpurun_out/pytest_refpy.log
