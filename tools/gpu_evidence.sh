#!/bin/bash
# One GPU session that refreshes every piece of committed evidence (about 4 minutes on a B200): the full GPU test suite,
# smoke, benches (batch 1 serial / 6 lanes, batch 32, batch 8), the reference port on the GPU, the K1 sweep of BASELINE
# configs[4], the renderer bench, the ncu launch list and full captures of K3 / back-projection / renderer kernels.
# Outputs in gpurun_out/ (kept far below gpurun's 64 MiB limit: never capture --set full of every kernel).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
. tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
stamp "smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
stamp "bench"
run_bench b1_l1 X=1 -- --steps 200 --warmup 20 --lanes 1
run_bench b1_l6 X=1 -- --steps 200 --warmup 20 --lanes 6
run_bench b32_l1 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
run_bench b8_l2 X=1 -- --steps 40 --warmup 5 --batch 8 --lanes 2
stamp "reference port on the GPU"; timeout 600 python tools/ref_gpu_port.py --steps 20 --warmup 5 --check > gpurun_out/ref_gpu_port.txt 2>&1; tail -2 gpurun_out/ref_gpu_port.txt
stamp "k1 sweep"; timeout 300 python tools/k1_sweep.py > gpurun_out/k1_sweep.txt 2>&1; tail -12 gpurun_out/k1_sweep.txt
stamp "render bench"; timeout 300 python tools/render_bench.py > gpurun_out/render_bench.json 2> gpurun_out/render_bench.err; cat gpurun_out/render_bench.json; tail -2 gpurun_out/render_bench.err
stamp "ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
stamp "ncu full + source, K3 at batch 8"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_jln' -s 1 -c 1 -o gpurun_out/prof_k3_b8 python tools/profile_driver.py 2 8 > gpurun_out/ncu_k3.log 2>&1; tail -1 gpurun_out/ncu_k3.log
stamp "ncu full, back-projection kernels at batch 1"
timeout 300 ncu --set full --clock-control none -k regex:'k3_jln|k1_hdn|k0_stage' -s 3 -c 3 -o gpurun_out/prof_bp_b1 python tools/profile_driver.py 2 1 > gpurun_out/ncu_bp.log 2>&1; tail -1 gpurun_out/ncu_bp.log
stamp "ncu full, renderer"
timeout 300 ncu --set full --clock-control none -k regex:'k_hm_render' -s 4 -c 1 -o gpurun_out/prof_render python tools/render_bench.py > gpurun_out/ncu_render.log 2>&1; tail -1 gpurun_out/ncu_render.log
stamp "official bench line (default flags)"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ gpurun_out/bench_default.json default
stamp done; du -sh gpurun_out
