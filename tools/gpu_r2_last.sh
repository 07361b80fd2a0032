mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --check > gpurun_out/bench_ours_check.json 2> gpurun_out/bench_ours_check.err; echo "check rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_ours_check.json')); print(d['value'], d['e2e']['value'], d.get('parity'))" | cut -c1-900
timeout 300 python tools/bench_configs.py cfg3 2>&1 | tail -1 > gpurun_out/cfg3.json; cat gpurun_out/cfg3.json | cut -c1-400
