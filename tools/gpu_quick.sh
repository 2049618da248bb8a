#!/bin/bash
# shortest useful GPU session: full GPU suite + the official bench line
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
stamp "official bench line (default flags)"
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ gpurun_out/bench_default.json default
stamp done
