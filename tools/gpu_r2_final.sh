#!/bin/bash
# round-2 final single-GPU visit: GPU suite, smoke, both bench arms, cfg4 / cfg5 lines, ncu launch list of the bench
# command, ncu --set full of the two compositing kernels, per-kernel metrics of one step (-> profiles/ncu_traffic.json)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:d[k] for k in ('value','ms_per_step','fwd_fps','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline'))
print({k:round(v,4) for k,v in d['kernels_ms'].items()})
for k,v in d['rooflines'].items():
    if not k.startswith('_'): print(k, v['bound'], round(v['achieved'],2), v['unit'], 'frac', round(v['frac'],4), 'traffic/alg', v.get('traffic_over_algorithmic'))
print(d['rooflines'].get('_ncu_capture'))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-600 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_n1.json')); print({k:d[k] for k in ('value','ms_per_step','scaling')}, d['e2e']['value'])"
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --check > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; echo "cfg5 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_n1.json')); print({k:d[k] for k in ('value','ms_per_step','scaling')}, d['e2e']['value'], {k:d['parity'].get(k) for k in ('pass','image_max_abs','rect_mismatch','error')} if d.get('parity') else None)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
bash tools/gpu_r2_ncu_comp.sh
bash tools/gpu_r2_ncu_step.sh > /dev/null 2>&1; ls -la gpurun_out/r2_step_metrics.csv
du -sh gpurun_out
