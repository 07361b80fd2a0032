#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rgb_path.py -q -s -m gpu 2>&1 | grep -E "csr|passed|failed|skipped|Error" | tail -8
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
