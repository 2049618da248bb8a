#!/bin/bash
# K3 pair-form A/B (FVP_K3_PAIR: 0 patch kernel, 1 pair form + shifted second heat-map copy, 2 pair form without the copy,
# 5 = 1 at 5 CTAs per SM): full GPU suite with the default, the parity files again under the pair form, stage times per
# mode at batch 1 and batch 32, the official bench line, one ncu --set full capture of the pair kernel.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu (library default)"; timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -6
stamp "parity files, FVP_K3_PAIR=1"
FVP_K3_PAIR=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -s > gpurun_out/k3pair_pytest_1.log 2>&1; tail -12 gpurun_out/k3pair_pytest_1.log
for m in 0 1 5 2; do stamp "batch 1, FVP_K3_PAIR=$m"; run_bench k3pair${m}_b1 FVP_K3_PAIR=$m -- --steps 200 --warmup 20; done
for m in 0 1 5; do stamp "batch 32, FVP_K3_PAIR=$m"; run_bench k3pair${m}_b32 FVP_K3_PAIR=$m -- --steps 20 --warmup 5 --batch 32 --lanes 1; done
stamp "parity file, FVP_K3_PAIR=5 / 2"
FVP_K3_PAIR=5 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/k3pair_pytest_5.log 2>&1; tail -3 gpurun_out/k3pair_pytest_5.log
FVP_K3_PAIR=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/k3pair_pytest_2.log 2>&1; tail -3 gpurun_out/k3pair_pytest_2.log
stamp "ncu full, pair kernel + K0 at batch 1"
FVP_K3_PAIR=1 timeout 200 ncu --set full --clock-control none -k regex:'k3_jln|k0_stage' -s 3 -c 2 -f -o gpurun_out/r02_prof_k3pair_b1 python tools/profile_driver.py 2 1 > gpurun_out/ncu_k3pair.log 2>&1; tail -1 gpurun_out/ncu_k3pair.log
stamp "official bench line (default flags)"
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ gpurun_out/bench_default.json default
stamp done; du -sh gpurun_out
