#!/bin/bash
# GPU session 4: full test suite (renderer, config 5, split pose head, K3 with RED-folded planes), benches, K1 sweep,
# renderer bench, ncu launch list + captures of K3 / back-projection / renderer kernels.
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
