#!/usr/bin/env bash
# Builds libhpusm.so (all hand-written sm_100a kernels + the C ABI) in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libhpusm.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -cudart static)
OBJS=()
PIDS=()
for f in *.cu; do
  o="build/${f%.cu}.o"
  mkdir -p build
  stale=0
  for h in *.cuh *.inc ../../include/hpusm.h; do [[ "$h" -nt "$o" ]] && stale=1; done
  if [[ ! -f "$o" || "$f" -nt "$o" || $stale -eq 1 ]]; then
    echo "nvcc $f"
    "$NVCC" "${FLAGS[@]}" ${HPUSM_PTXAS_V:+-Xptxas -v} ${HPUSM_EXTRA_FLAGS:-} -c "$f" -o "$o" &
    PIDS+=($!)
  fi
  OBJS+=("$o")
done
for p in "${PIDS[@]:-}"; do   # a failed compile must fail the build, not link a stale object
  if [[ -n "$p" ]] && ! wait "$p"; then echo "build.sh: a compile failed" >&2; rm -f "$OUT"; exit 1; fi
done
"$NVCC" -shared -cudart static "${OBJS[@]}" -o "$OUT" -lpthread -ldl
echo "built $(readlink -f $OUT)"
