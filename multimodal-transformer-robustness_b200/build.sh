#!/usr/bin/env bash
# Build libmultb200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/lib"
# MTB_VARIANT=trace builds lib/libmultb200_trace.so with in-kernel timestamps (tools/tc_trace.py)
VARIANT="${MTB_VARIANT:-}"
BUILD="$HERE/build${VARIANT:+/$VARIANT}"
LIBNAME="libmultb200${VARIANT:+_$VARIANT}.so"
EXTRA=()
[[ "$VARIANT" == "trace" ]] && EXTRA+=(-DMTB_TC_TRACE)
[[ "$VARIANT" == "tracefine" ]] && EXTRA+=(-DMTB_TC_TRACE -DMTB_TC_TRACE_FINE)
[[ "$VARIANT" == "pdlearly" ]] && EXTRA+=(-DMTB_PDL_EARLY_WAIT)       # A/B: griddepcontrol.wait back at the top of the tcgen05 kernels
mkdir -p "$OUT" "$BUILD"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v "${EXTRA[@]}")
objs=()
pids=()
for f in "$HERE"/csrc/*.cu; do
  o="$BUILD/$(basename "${f%.cu}").o"
  objs+=("$o")
  if [[ ! -f "$o" || "$f" -nt "$o" || "$HERE/csrc/common.cuh" -nt "$o" || "$HERE/../include/multb200.h" -nt "$o" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/$LIBNAME" "${objs[@]}"
echo "built $OUT/$LIBNAME"
