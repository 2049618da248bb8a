#!/bin/bash
# GPU session 2: parity of the K3 patch kernel and the engine lanes, A/B benches (lanes 1..4, K3 v1/v2/occupancy),
# the reference-port GPU baseline, small ncu captures.  Outputs in gpurun_out/ (kept far below the 64 MiB limit).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
stamp() { echo "== $1 [$(( $(date +%s) - T0 )) s]"; }
summ() {  # summ <json file> <tag>
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    st = d['kernels']['stage_ms']
    print("[%s] FPS %.1f (%.3f ms/step) e2e %.1f serial %.1f | k3 %.3f p2p %.3f cn %.3f c2c %.3f pose %.3f k0+k1 %.3f | roofline %.1f GB/s" % (
        sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('serial_fps', 0), st['k3_jln_project'], st['p2p_net'],
        st['center_net'], st['proposals_c2c'], st['pose_head'], st['k0_stage'] + st['k1_hdn_project'], d['roofline']['achieved']))
except Exception as e:
    print("[%s] bench failed: %s" % (sys.argv[2], e)); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
}
run_bench() {  # run_bench <tag> <env...> -- <args...>
  tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  summ gpurun_out/bench_$tag.json $tag
}
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
