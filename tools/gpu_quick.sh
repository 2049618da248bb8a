#!/bin/bash
# quick GPU session: tests + bench (b1, b32) + launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench b1"; timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json')); print('FPS',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'cpu',d.get('cpu_baseline')); print(d['kernels']['stage_ms']); print('roofline',d['roofline']['achieved'],d['roofline']['frac'], 'k1', d['kernels']['k0+k1_hdn_project'])
except Exception as e: print('bench failed',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
echo "== bench b32"; timeout 600 python bench.py --steps 20 --warmup 5 --batch 32 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_b32.json')); print('FPS',round(d['value'],1),'ms',round(d['ms_per_step'],3)); print(d['kernels']['stage_ms'])
except Exception as e: print('bench failed',e); print(open('gpurun_out/bench_b32.err').read()[-2000:])
PY
