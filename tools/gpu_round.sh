#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list, ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
cat gpurun_out/bench_ours.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'composite_(fwd|bwd)_kernel' -s 6 -c 2 -o gpurun_out/prof_composite -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_comp.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
