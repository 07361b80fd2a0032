#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m pytest tests/test_gpu_parallel.py -q -s -k "tile_sharded or bands" 2>&1 | grep -E "passed|failed|Error|assert" | tail -5
for g in root allgather; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --band-gather $g > gpurun_out/bench_cfg5_${g}_n$N.json 2> gpurun_out/bench_cfg5_${g}_n$N.err; echo "cfg5 $g rc=$?"
python - "$g" "$N" <<'PY'
import json,sys
try:
    txt=open(f'gpurun_out/bench_cfg5_{sys.argv[1]}_n{sys.argv[2]}.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','scaling','sharded_vs_single_gpu_max_abs')}, 'e2e', d['e2e']['value'])
except Exception as e: print("no json:", e); print(open(f'gpurun_out/bench_cfg5_{sys.argv[1]}_n{sys.argv[2]}.err').read()[-1500:])
PY
done
