#!/bin/bash
# final visit of a round: all GPU tests, smoke, our bench arm (with cpu_baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/bench_ours.json
