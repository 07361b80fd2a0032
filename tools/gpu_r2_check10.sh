#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/bench_composite.py cfg3 10 | tail -1; echo "cfg3 rc=$?"
timeout 120 python tools/bench_composite.py cfg2 10 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python tools/profile_step.py cfg3 2>&1 | grep -v -i warn | sed -n 1,4p
