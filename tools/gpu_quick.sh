#!/bin/bash
# shortest useful GPU session: full GPU suite + the official bench line (+ what the box says about its host topology,
# + the Campus / Shelf lines if time is left)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
stamp "official bench line (default flags)"
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ gpurun_out/bench_default.json default
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("gpu_reference_port:", json.dumps(d.get("gpu_reference_port")), "vs:", d.get("vs_gpu_reference_port"))
PY
stamp "host topology"
PYTHONPATH=faster-voxelpose_b200 python - <<'PY'
import os, torch
from fvp import dist as D
print("cpus allowed", len(os.sched_getaffinity(0)), "numa nodes online", open("/sys/devices/system/node/online").read().strip(),
      "gpu nodes", D.local_gpu_numa_nodes(torch.cuda.device_count()),
      "plan for 2 ranks", [len(p) for p in D.plan_rank_cores(sorted(os.sched_getaffinity(0)), 2, D.local_gpu_numa_nodes(1) * 2)])
PY
for ps in campus shelf; do stamp "bench --preset $ps"; run_bench $ps X=1 -- --preset $ps --steps 100 --warmup 10; done
stamp done
