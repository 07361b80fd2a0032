#!/bin/bash
# fresh ncu --set full of the shipped compositing kernels (one launch each, cfg 2)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'composite_fwd' -s 2 -c 1 -o gpurun_out/prof_r2_fwd -f python tools/bench_composite.py cfg2 2 > gpurun_out/ncu_r2_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'composite_bwd' -s 2 -c 1 -o gpurun_out/prof_r2_bwd -f python tools/bench_composite.py cfg2 2 > gpurun_out/ncu_r2_bwd.log 2>&1; echo "ncu bwd rc=$?"
ls -la gpurun_out/prof_r2_*.ncu-rep
timeout 200 python tools/bench_composite.py cfg2 10 2>&1 | tail -1
