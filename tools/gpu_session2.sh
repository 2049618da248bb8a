#!/bin/bash
# GPU session 2: parity of the K3 patch kernel and the engine lanes, A/B benches (lanes 1..4, K3 v1/v2/occupancy),
# the reference-port GPU baseline, small ncu captures.  Outputs in gpurun_out/ (kept far below the 64 MiB limit).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
. tools/gpu_lib.sh
stamp "pytest -m gpu (K3 patch kernel default)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
stamp "pytest K3 slab kernel"; FVP_K3_VERSION=1 timeout 600 python -m pytest tests -m gpu -x -q -k "k3 or end_to_end or empty_and_border" 2>&1 | tail -3
stamp "bench b1 lanes"
for l in 1 2 3 4; do run_bench b1_l$l X=1 -- --steps 200 --warmup 20 --lanes $l; done
stamp "bench b1 K3 variants (lanes 1)"
run_bench b1_k3v1 FVP_K3_VERSION=1 -- --steps 200 --warmup 20 --lanes 1
run_bench b1_k3occ3 FVP_K3_OCC=3 -- --steps 200 --warmup 20 --lanes 1
stamp "bench b32"
run_bench b32_l1 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
run_bench b32_l2 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 2
run_bench b32_k3v1 FVP_K3_VERSION=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
run_bench b32_k3occ3 FVP_K3_OCC=3 -- --steps 20 --warmup 5 --batch 32 --lanes 1
run_bench b8_l2 X=1 -- --steps 40 --warmup 5 --batch 8 --lanes 2
stamp "reference port on the GPU"; timeout 600 python tools/ref_gpu_port.py --steps 20 --warmup 5 --check > gpurun_out/ref_gpu_port.txt 2>&1; tail -3 gpurun_out/ref_gpu_port.txt
stamp "ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
stamp "ncu full + source, K3 at batch 8"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_jln' -s 1 -c 1 -o gpurun_out/prof_k3_b8 python tools/profile_driver.py 2 8 > gpurun_out/ncu_k3.log 2>&1; tail -1 gpurun_out/ncu_k3.log
stamp "ncu full, back-projection kernels at batch 1"
timeout 300 ncu --set full --clock-control none -k regex:'k3_jln|k3b|k1_hdn|k0_stage' -s 4 -c 4 -o gpurun_out/prof_bp_b1 python tools/profile_driver.py 2 1 > gpurun_out/ncu_bp.log 2>&1; tail -1 gpurun_out/ncu_bp.log
stamp "official bench line (default flags)"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ gpurun_out/bench_default.json default; tail -c 400 gpurun_out/bench_default.json
stamp done; du -sh gpurun_out; ls -la gpurun_out
