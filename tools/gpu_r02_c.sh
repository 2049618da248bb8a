#!/bin/bash
# round-2 session C: the TMA-fed split-activation conv path - parity, per-layer A/B against the legacy loaders, bench A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "conv parity (incl. debug mode 4 = TMA path)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "conv or p2p or center_net" 2>&1 | tail -15
stamp "layers legacy (mode 2)"; FVP_CONV_DEBUG_MODE=2 timeout 200 python tools/conv_layers.py 30 960 2>&1 | tail -21
stamp "layers TMA (mode 4)"; FVP_CONV_DEBUG_MODE=4 timeout 200 python tools/conv_layers.py 30 960 2>&1 | tail -21
stamp "full gpu suite"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
stamp "bench, split activations OFF"; run_bench r02_c_legacy FVP_SPLIT_ACT=0 -- --steps 200 --warmup 20
stamp "bench, split activations ON"; run_bench r02_c_split FVP_SPLIT_ACT=1 -- --steps 200 --warmup 20
stamp "bench b32, split OFF"; run_bench r02_c_b32_legacy FVP_SPLIT_ACT=0 -- --steps 20 --warmup 5 --batch 32 --lanes 1
stamp "bench b32, split ON"; run_bench r02_c_b32_split FVP_SPLIT_ACT=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
stamp done
