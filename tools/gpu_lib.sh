#!/bin/bash
# helpers shared by the GPU session scripts: stamp <label>, summ <bench json> <tag>, run_bench <tag> <ENV=..>... -- <bench args>
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
