#!/bin/bash
# Final evidence session of a round, most important first (the GPU budget may cut the tail): full GPU suite, smoke, the
# official bench line, batch-32 and serial batch-1 benches, ncu launch list, per-layer conv times.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8
stamp "smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
stamp "official bench line (default flags)"
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ gpurun_out/bench_default.json default
stamp "bench"
run_bench b32_l1 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
run_bench b1_l1 X=1 -- --steps 200 --warmup 20 --lanes 1
stamp "conv layers"; timeout 120 python tools/conv_layers.py 30 960 > gpurun_out/conv_layers.txt 2>&1; tail -21 gpurun_out/conv_layers.txt
stamp "ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
stamp done; du -sh gpurun_out
