#!/bin/bash
# one 8-GPU bench run (default exchange = sparse NVLS pull)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_pull.json 2> gpurun_out/bench_n8_pull.err
echo "n8 rc=$?"; grep -v "^NCCL\|^\*\|^Setting" gpurun_out/bench_n8_pull.json | tail -1 | cut -c1-220; grep -o '"kernels_ms.*"roofline' gpurun_out/bench_n8_pull.json | cut -c1-700
