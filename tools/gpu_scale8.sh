#!/bin/bash
# 8-GPU visit: bench.py at N=8 (all exchange modes) and N=4 (push), cfg4 strong-scaling step, cfg5 tile-sharded frame
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for mode in push sparse allreduce fused; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --dp-mode $mode > gpurun_out/bench_n8_$mode.json 2> gpurun_out/bench_n8_$mode.err
  echo "n8 $mode rc=$?"; tail -1 gpurun_out/bench_n8_$mode.json | cut -c1-200; grep -o '"kernels_ms.*"roofline' gpurun_out/bench_n8_$mode.json | cut -c1-400
done
timeout 300 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4_push.json 2> gpurun_out/bench_n4_push.err; echo "n4 rc=$?"; tail -1 gpurun_out/bench_n4_push.json | cut -c1-200
timeout 300 $TR --nproc-per-node 8 --master-port 29523 tools/bench_configs.py cfg4 10 pull 2> gpurun_out/cfg4_n8.err | tee gpurun_out/cfg4_n8.json
timeout 300 $TR --nproc-per-node 8 --master-port 29524 tools/bench_configs.py cfg5 10 2> gpurun_out/cfg5_n8.err | tee gpurun_out/cfg5_n8.json
