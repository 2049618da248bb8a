#!/bin/bash
# Build libfvp_b200.so for sm_100a (cross-compiles without a GPU).  Usage: csrc/build.sh [outdir]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${1:-$HERE/..}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off
       -Xptxas -v)
# experiments: extra -D switches (e.g. FVP_NVCC_DEFS="-DFVP_K3_PREFETCH_GRID" csrc/build.sh); empty in every shipped build
if [ -n "${FVP_NVCC_DEFS:-}" ]; then read -r -a EXTRA <<< "$FVP_NVCC_DEFS"; FLAGS+=("${EXTRA[@]}"); fi
SRCS=(fvp_api.cu fvp_params.cu fvp_backproject.cu fvp_conv.cu fvp_conv_tc.cu fvp_proposal.cu fvp_pose.cu fvp_render.cu fvp_backbone.cu)
mkdir -p "$HERE/build"
pids=()
for s in "${SRCS[@]}"; do
  ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$s" -o "$HERE/build/${s%.cu}.o" > "$HERE/build/${s%.cu}.log" 2>&1 ) &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
if [ "$fail" != 0 ]; then cat "$HERE"/build/*.log | grep -E "error|Error" -B2 -A4 >&2 || true; exit 1; fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libfvp_b200.so" "$HERE"/build/*.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT/libfvp_b200.so"
