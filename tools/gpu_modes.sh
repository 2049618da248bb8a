#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python tools/tc_check.py ${1:-panoptic_256x192} > gpurun_out/tc.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/tc.log
for m in 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --batch 32 --no-cpu-baseline --conv-mode $m > gpurun_out/b32_m$m.json 2>gpurun_out/b32_m$m.err
  python - $m <<'PY'
import sys,json
m=sys.argv[1]
try:
    d=json.load(open('gpurun_out/b32_m%s.json'%m)); print("b32 mode",m,round(d["value"],1),{k:round(v,3) for k,v in d["kernels"]["stage_ms"].items()})
except Exception as e: print("fail",e, open('gpurun_out/b32_m%s.err'%m).read()[-1500:])
PY
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --conv-mode $m > gpurun_out/b1_m$m.json 2>gpurun_out/b1_m$m.err
  python - $m <<'PY'
import sys,json
m=sys.argv[1]
try:
    d=json.load(open('gpurun_out/b1_m%s.json'%m)); print("b1 mode",m,round(d["value"],1),{k:round(v,3) for k,v in d["kernels"]["stage_ms"].items()})
except Exception as e: print("fail",e, open('gpurun_out/b1_m%s.err'%m).read()[-1500:])
PY
done
