#!/bin/bash
# round-2 session D: backbone slice test (N2), multi-rank test, smoke
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; . tools/gpu_lib.sh
stamp "backbone slice + multirank"; timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_multirank.py -q -x -s 2>&1 | tail -25
stamp "smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
stamp done
