#!/bin/bash
# per-kernel DRAM bytes + utilisation of ONE eager cfg-2 step (all kernels), as a CSV small enough to travel back
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,lts__t_sector_hit_rate.pct
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --page raw --log-file gpurun_out/r2_step_metrics.csv python tools/one_step.py > gpurun_out/ncu_r2_step.log 2>&1; echo "ncu step rc=$?"; tail -2 gpurun_out/ncu_r2_step.log
ls -la gpurun_out/r2_step_metrics.csv; head -c 600 gpurun_out/r2_step_metrics.csv
