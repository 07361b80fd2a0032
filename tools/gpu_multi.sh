#!/bin/bash
# multi-GPU visit: parallel GPU tests + bench at N ranks (both exchange modes).  usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q -rA > gpurun_out/pytest_gpu_parallel.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu_parallel.log
for mode in ${MODES:-push fused sparse allreduce}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --dp-mode $mode > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  echo "bench $mode rc=$?"; tail -1 gpurun_out/bench_n${N}_$mode.json | cut -c1-1500; tail -3 gpurun_out/bench_n${N}_$mode.err
done
