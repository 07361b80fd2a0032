#!/bin/bash
# multi-GPU visit (gpurun --gpus N): 2-GPU tests, default bench, cfg4, cfg5 at N ranks
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests/test_gpu_parallel.py -x -q -s > gpurun_out/pytest_gpu_parallel.log 2>&1; echo "pytest parallel rc=$?"
grep -E "passed|failed|dp rank" gpurun_out/pytest_gpu_parallel.log | tail -20
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 $2 > gpurun_out/bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err; echo "$1 rc=$?"
  tail -2 gpurun_out/bench_$1_n$N.err | cut -c1-300
  python - "$1" "$N" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}_n{sys.argv[2]}.json'))
    keys=('value','ms_per_step','scaling','exchange_check','exchange_union_rows','exchange_bytes_ingested_per_gpu','sharded_vs_single_gpu_max_abs','rank_kernel_ms')
    print({k:d.get(k) for k in keys}, d['e2e']['value'])
    print({k:round(v,4) for k,v in (d.get('kernels_ms') or {}).items()})
except Exception as e: print("no json:", e)
PY
}
run default ""
run cfg4 "--workload cfg4"
STEPS=10 run cfg5 "--workload cfg5"
