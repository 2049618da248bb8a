#!/bin/bash
# ncu --set full + source of k_conv_tc on single layers (steady state, n = 960 images)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "3x3 32->32"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 1 -c 1 -f -o gpurun_out/prof_conv_32 python tools/conv_ncu.py 960 64 64 32 32 3 2>&1 | tail -2
stamp "3x3 64->64"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 1 -c 1 -f -o gpurun_out/prof_conv_64 python tools/conv_ncu.py 960 32 32 64 64 3 2>&1 | tail -2
stamp done; ls -la gpurun_out/*.ncu-rep
