#!/bin/bash
# A/B of prebuilt library variants under ab_libs/<tag>/libfvp_b200.so (FVP_B200_LIB selects the library): per-layer conv times
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
for tag in "$@"; do
  stamp "conv layers [$tag]"
  FVP_B200_LIB=$PWD/ab_libs/$tag/libfvp_b200.so FVP_B200_LAX_SYMBOLS=1 timeout 300 python tools/conv_layers.py 30 960 2>&1 | tail -21
done
stamp done
