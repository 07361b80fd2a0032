#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_ours.err
python -c "
import json; d=json.load(open('gpurun_out/bench_ours.json')); print({k:d[k] for k in ('value','ms_per_step','fwd_fps')}, 'e2e', d['e2e']['value'])"
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_n1.json')); print({k:d[k] for k in ('value','ms_per_step','scaling')}, d['e2e']['value'])"
