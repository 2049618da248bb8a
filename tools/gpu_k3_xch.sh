#!/bin/bash
# K3 tap-exchange A/B (FVP_K3_XCH: 0 five shuffles, 1 three shuffles + rebuilt weights, 2 shared-memory exchange)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
for m in 1 2; do stamp "parity file, FVP_K3_XCH=$m"; FVP_K3_XCH=$m timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3; done
for m in 0 1 2; do stamp "batch 32, FVP_K3_XCH=$m"; run_bench k3xch${m}_b32 FVP_K3_XCH=$m -- --steps 20 --warmup 5 --batch 32 --lanes 1; done
for m in 0 1 2; do stamp "batch 1, FVP_K3_XCH=$m"; run_bench k3xch${m}_b1 FVP_K3_XCH=$m -- --steps 200 --warmup 20; done
stamp done
