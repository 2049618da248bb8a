#!/bin/bash
# Round-2 A/B of the compile-time-gated K3 experiments (csrc/fvp_backproject.cu): rebuilds the library on the GPU box for
# every switch combination, runs the K3 / end-to-end parity tests and a batch-32 + batch-1 bench, restores the default build.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
for defs in "" "-DFVP_K3_PREFETCH_GRID" "-DFVP_K3_SPLIT_BARRIER" "-DFVP_K3_PREFETCH_GRID -DFVP_K3_SPLIT_BARRIER" "-DFVP_K3_INKERNEL_PROJ"; do
  tag=$(echo "default$defs" | tr -d ' ' | sed 's/-DFVP_K3_/_/g')
  stamp "build [$defs]"; FVP_NVCC_DEFS="$defs" bash faster-voxelpose_b200/csrc/build.sh | tail -1
  stamp "parity [$defs]"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "k3 or end_to_end or crops or determinism" 2>&1 | tail -2
  run_bench b32_$tag X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
  run_bench b1_$tag X=1 -- --steps 200 --warmup 20 --lanes 6
done
stamp "restore default build"; bash faster-voxelpose_b200/csrc/build.sh | tail -1
stamp done
