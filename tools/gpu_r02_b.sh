#!/bin/bash
# round-2 session B: proposal / conv parity after the C2C weight ring and the conv codegen fixes, conv A/B, bench
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for tag in cs0 cs1 cs2; do
  stamp "conv layers [$tag]"
  FVP_B200_LIB=$PWD/ab_libs/$tag/libfvp_b200.so timeout 300 python tools/conv_layers.py 30 960 2>&1 | grep -E "7x7|3x3 16->32|3x3 32->32|3x3 64->64|128->128|sum"
done
stamp "official bench line (default flags)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; summ gpurun_out/r02_bench_b.json default
stamp "bench batch 32 serial"; run_bench r02_b32 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
stamp done
