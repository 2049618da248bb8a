#!/bin/bash
# conv A/B session: per-layer steady-state times with value checks, benches; OCCS = values of FVP_TC_OCC to try
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
for o in ${OCCS:-2}; do
  stamp "layers occ=$o"; FVP_TC_OCC=$o timeout 200 python tools/conv_layers.py ${NS:-960} 2>&1 | tail -12
  stamp "bench occ=$o"
  run_bench b32_occ$o FVP_TC_OCC=$o -- --steps 20 --warmup 5 --batch 32 --lanes 1
  [ -n "$B1" ] && run_bench b1_occ$o FVP_TC_OCC=$o -- --steps 200 --warmup 20 --lanes 6
done
stamp done
