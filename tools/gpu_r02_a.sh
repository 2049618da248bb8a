#!/bin/bash
# round-2 session A: full GPU suite (with the noise-floor prints), smoke, official bench line, Campus / Shelf bench lines,
# N-tile variant sweep of the conv planner at 30 images
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log; grep "noise floor" gpurun_out/r02_pytest_gpu.log
stamp "smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
stamp "official bench line (default flags)"
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; summ gpurun_out/r02_bench_default.json default
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_default.json"))
    for k in ("plugin_forward", "gpu_reference_port", "vs_gpu_reference_port", "cpu_baseline", "timed_reps", "timed_total_ms", "serial_ms_p50"):
        print(k, d.get(k))
    print("e2e", d["e2e"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02_bench_default.err").read()[-3000:])
PY
for ps in campus shelf; do
  stamp "bench --preset $ps"; run_bench r02_$ps X=1 -- --preset $ps --steps 100 --warmup 10
  mv gpurun_out/bench_r02_$ps.json gpurun_out/r02_bench_$ps.json 2>/dev/null
done
stamp "bench batch 32 serial"; run_bench r02_b32 X=1 -- --steps 20 --warmup 5 --batch 32 --lanes 1
for v in 0 1 2; do stamp "conv layers, FVP_TC_VARIANT=$v, n=30"; FVP_TC_VARIANT=$v timeout 200 python tools/conv_layers.py 30 2>&1 | tail -10; done
stamp done
