#!/bin/bash
# round-2 evidence session: GPU suite, bench lines (default / batch 32 / serial / Campus / Shelf),
# ncu launch list of one frame, ncu --set full of the back-projection kernels and of the TMA-fed 3x3 32->32 conv layer
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
stamp "official bench line (default flags)"
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; summ gpurun_out/r02_bench_default.json default
stamp "bench batch 32"; run_bench r02_b32 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
stamp "bench serial (1 lane)"; run_bench r02_b1_l1 X=1 -- --steps 200 --warmup 20 --lanes 1
for ps in campus shelf; do stamp "bench --preset $ps"; run_bench r02_$ps X=1 -- --preset $ps --steps 100 --warmup 10; done
stamp "ncu launch list (one eager frame, batch 1)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
stamp "ncu launch list, Campus (J = 17)"
FVP_PROFILE_PRESET=campus timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_campus.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list_campus.log 2>&1; tail -1 gpurun_out/ncu_list_campus.log
stamp "ncu full, back-projection kernels at batch 1"
timeout 300 ncu --set full --clock-control none -k regex:'k3_jln|k1_hdn|k0_stage' -s 3 -c 3 -f -o gpurun_out/r02_prof_bp_b1 python tools/profile_driver.py 2 1 > gpurun_out/ncu_bp.log 2>&1; tail -1 gpurun_out/ncu_bp.log
stamp "ncu full + source, TMA-fed 3x3 32->32 at 960 images"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 1 -c 1 -f -o gpurun_out/r02_prof_conv_32_tma python tools/conv_ncu.py 960 64 64 32 32 3 4 2>&1 | tail -2
stamp "ncu full + source, proposal stage (8-CTA cluster per column) at batch 1"
timeout 300 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:k_proposals_cluster --launch-skip 2 -c 1 -f -o gpurun_out/r02_prof_c2c_cluster python tools/profile_driver.py 4 1 2>&1 | tail -1
stamp "backbone (N2): 5 views of 960x512"; timeout 300 python tools/backbone_bench.py 50 5 512 960 10 2>&1 | tail -1
stamp "ncu full, TMA-fed 3x3 64->64 at 960 images"
timeout 300 ncu --set full --clock-control none -k regex:k_conv_tc -s 1 -c 1 -f -o gpurun_out/r02_prof_conv_64_tma python tools/conv_ncu.py 960 32 32 64 64 3 4 2>&1 | tail -2
stamp done; du -sh gpurun_out
