#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'composite_(fwd|bwd)_kernel' -s 6 -c 2 -o gpurun_out/prof_composite -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_comp.log 2>&1; echo "ncu rc=$?"
