#!/bin/bash
# quick conv session: parity tests that touch the conv engines, per-layer times with value checks, benches
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "conv tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "conv or p2p or center_net or end_to_end or determinism or lanes" 2>&1 | tail -3
stamp "layers"; timeout 120 python tools/conv_layers.py ${NS:-30 960} 2>&1 | tail -21
stamp "bench"
run_bench b1_l1 X=1 -- --steps 200 --warmup 20 --lanes 1
run_bench b32_l1 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
stamp done
