#!/bin/bash
# GPU visit: all GPU tests, our bench arm, cfg3 full-training-step side bench (optimiser variants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 600 python tools/bench_configs.py cfg3 20 > gpurun_out/cfg3.json 2> gpurun_out/cfg3.err; echo "cfg3 rc=$?"
cat gpurun_out/cfg3.json; tail -3 gpurun_out/cfg3.err
