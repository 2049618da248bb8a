#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list + full captures. Outputs in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench batch 32"; timeout 600 python bench.py --steps 20 --warmup 5 --batch 32 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; tail -c 1500 gpurun_out/bench_b32.json; tail -3 gpurun_out/bench_b32.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_driver.py 2 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_jln_project|k1_hdn|k_conv_nhwc|k_proposals|k_pose_head|k0_stage' -s 56 -c 40 -o gpurun_out/prof_full python tools/profile_driver.py 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
