#!/bin/bash
# conv check + ncu --set full (with source) of k_conv_tc on the 3x3 32->32 layer at 960 images (two CTAs per SM)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "conv tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "conv or p2p or center_net or end_to_end" 2>&1 | tail -3
stamp "layers"; timeout 120 python tools/conv_layers.py 960 2>&1 | tail -10
stamp "ncu 3x3 32->32"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 1 -c 1 -f -o gpurun_out/prof_conv_32_occ2 python tools/conv_ncu.py 960 64 64 32 32 3 2>&1 | tail -2
stamp done
