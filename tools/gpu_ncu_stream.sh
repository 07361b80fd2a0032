#!/bin/bash
# ncu --set full of the HBM-bound streaming kernels of one cfg-2 step (K1, K2's passes, K4b)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'project_cull_fused_kernel|radix_scatter_kernel|radix_hist_kernel|emit_kernel|project_backward_kernel|ranges_kernel|count_sorted_kernel' \
  -s 150 -c 17 -o gpurun_out/prof_stream -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stream.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_stream.log
ls -la gpurun_out/prof_stream.ncu-rep
