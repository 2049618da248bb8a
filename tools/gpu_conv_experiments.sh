#!/bin/bash
# Round-2 A/B of the compile-time-gated conv experiments: (a) 16 instead of 8 loader warps in the one-CTA-per-SM variant
# (-DFVP_TC_LOADERS=512; aimed at the loader-bound 64-channel layers).  Compare the per-layer times of the 64- and
# 128-channel rows; the full-resolution rows lose their second CTA in this build and are expected to be slower;
# (b) -DFVP_TC_STREAM_ROWS: streamed weights move one tap ROW per ring stage (one barrier wait / commit per 3 taps):
# compare the 128-channel rows at 960 images.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
for defs in "" "-DFVP_TC_LOADERS=512" "-DFVP_TC_STREAM_ROWS"; do
  stamp "build [$defs]"; FVP_NVCC_DEFS="$defs" bash faster-voxelpose_b200/csrc/build.sh | tail -1
  stamp "conv parity [$defs]"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "conv or p2p or center_net" 2>&1 | tail -2
  stamp "layers [$defs]"; timeout 200 python tools/conv_layers.py 30 960 2>&1 | tail -21
done
stamp "restore default build"; bash faster-voxelpose_b200/csrc/build.sh | tail -1
stamp done
