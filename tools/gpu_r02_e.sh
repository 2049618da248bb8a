#!/bin/bash
# round-2 session E: per-launch times of one batch-32 frame set (where does P2PNet's time go at large batch), memcheck of the conv paths
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "ncu launch list, batch 32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b32.csv python tools/profile_driver.py 2 32 > gpurun_out/ncu_list_b32.log 2>&1; tail -1 gpurun_out/ncu_list_b32.log
stamp "compute-sanitizer memcheck: single conv layers (legacy + TMA) and one full forward"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "single_conv or live_oracle" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_memcheck.log
stamp done
