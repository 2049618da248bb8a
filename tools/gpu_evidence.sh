#!/bin/bash
# One GPU session that refreshes every piece of committed evidence: tests, smoke, bench (batch 1 with the CPU baseline,
# batch 32), ncu launch list, ncu --set full of one whole forward (no source) and of K3/K1 (with source), conv role
# counters.  Outputs in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== smoke [$(( $(date +%s) - T0 )) s]"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench b1 [$(( $(date +%s) - T0 )) s]"; timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
tail -c 2500 gpurun_out/bench_b1.json; tail -2 gpurun_out/bench_b1.err
echo "== bench b32 [$(( $(date +%s) - T0 )) s]"; timeout 600 python bench.py --steps 20 --warmup 5 --batch 32 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err
tail -c 1800 gpurun_out/bench_b32.json; tail -2 gpurun_out/bench_b32.err
echo "== ncu launch list [$(( $(date +%s) - T0 )) s]"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
echo "== ncu full, every kernel of one forward [$(( $(date +%s) - T0 )) s]"
timeout 900 ncu --set full --clock-control none -k regex:'k3_jln|k3b|k1_hdn|k0_stage|k_conv_tc|k_proposals|k_pose_head|k_nms|k_finalize|k_maxpool' -c 80 -o gpurun_out/prof_all python tools/profile_driver.py 1 > gpurun_out/ncu_all.log 2>&1; tail -1 gpurun_out/ncu_all.log
echo "== ncu full + source, K3 at batch 8 [$(( $(date +%s) - T0 )) s]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_jln' -s 1 -c 1 -o gpurun_out/prof_k3_b8 python tools/profile_driver.py 2 8 > gpurun_out/ncu_k3.log 2>&1; tail -1 gpurun_out/ncu_k3.log
echo "== conv roles [$(( $(date +%s) - T0 )) s]"; timeout 300 python tools/conv_roles.py > gpurun_out/conv_roles.txt 2>&1; tail -20 gpurun_out/conv_roles.txt
echo "== done [$(( $(date +%s) - T0 )) s]"; ls -la gpurun_out
