#!/bin/bash
# whole GPU suite on a 2-GPU box (single-GPU tests + the 2-rank exchange / ADC / band tests) + the 2-GPU bench lines
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"
grep -E "dp rank|passed|failed|Error" gpurun_out/pytest_gpu_2gpu.log | tail -20
timeout 300 python tools/bench_binning.py cfg2 --kernels 2>&1 | grep -v -i warn | tail -18
N=2
run() {
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 $2 > gpurun_out/bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err; echo "$1 rc=$?"
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
run cfg5 "--workload cfg5 --steps 5"
