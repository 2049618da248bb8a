#!/bin/bash
# GPU session 3: renderer + config-5 tests, lanes sweep, K3 after the divergence fixes, K1 sweep, renderer bench, ncu of K3.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
. tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
stamp "bench b1 lanes"
for l in 1 4 6 8; do run_bench b1_l$l X=1 -- --steps 200 --warmup 20 --lanes $l; done
stamp "bench b32"
run_bench b32_l1 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
stamp "k1 sweep"; timeout 300 python tools/k1_sweep.py > gpurun_out/k1_sweep.txt 2>&1; cat gpurun_out/k1_sweep.txt | tail -12
stamp "render bench"; timeout 300 python tools/render_bench.py > gpurun_out/render_bench.json 2> gpurun_out/render_bench.err; cat gpurun_out/render_bench.json; tail -2 gpurun_out/render_bench.err
stamp "ncu full + source, K3 at batch 8"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_jln' -s 1 -c 1 -o gpurun_out/prof_k3_b8 python tools/profile_driver.py 2 8 > gpurun_out/ncu_k3.log 2>&1; tail -1 gpurun_out/ncu_k3.log
stamp done; du -sh gpurun_out
