#!/bin/bash
# A/B session: tests, then bench b1/b32 under environment variants given as arguments ("VAR=1" strings, "" = default)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for variant in "" "$@"; do
  for b in 1 32; do
    steps=200; [ $b = 32 ] && steps=20
    env $variant timeout 600 python bench.py --steps $steps --warmup 10 --batch $b --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python - "$variant" $b <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/ab.json'))
    print(f"[{sys.argv[1] or 'default'}] b{sys.argv[2]} FPS {d['value']:.1f} ms {d['ms_per_step']:.3f} e2e {d['e2e']['value']:.1f}", {k: round(v, 3) for k, v in d['kernels']['stage_ms'].items()})
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/ab.err').read()[-1500:])
PY
  done
done
