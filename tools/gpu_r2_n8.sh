#!/bin/bash
# 8-GPU lines: view-sharded weak scaling (default), cfg4 strong scaling, cfg5 tile-sharded render
mkdir -p gpurun_out
N=${N:-8}
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline $2 > gpurun_out/bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err; echo "$1 rc=$?"
  python - "$1" "$N" <<'PY'
import json,sys
try:
    txt=open(f'gpurun_out/bench_{sys.argv[1]}_n{sys.argv[2]}.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    keys=('value','ms_per_step','scaling','exchange_check','exchange_union_rows','sharded_vs_single_gpu_max_abs','rank_kernel_ms')
    print({k:d.get(k) for k in keys}, d['e2e']['value'])
    print({k:round(v,4) for k,v in (d.get('kernels_ms') or {}).items()})
except Exception as e: print("no json:", e)
PY
}
run default ""
if [ -z "$ONLY_DEFAULT" ]; then
run cfg4 "--workload cfg4 --steps 10"
run cfg5 "--workload cfg5 --steps 5"
fi
