#!/usr/bin/env bash
# Builds libhpusm.so (all hand-written sm_100a kernels + the C ABI) in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libhpusm.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -cudart static)
OBJS=()
for f in *.cu; do
  o="build/${f%.cu}.o"
  mkdir -p build
  if [[ ! -f "$o" || "$f" -nt "$o" || common.cuh -nt "$o" || tc_ptx.cuh -nt "$o" || l2_internal.cuh -nt "$o" || ../../include/hpusm.h -nt "$o" ]]; then
    echo "nvcc $f"
    "$NVCC" "${FLAGS[@]}" ${HPUSM_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
  fi
  OBJS+=("$o")
done
wait
"$NVCC" -shared -cudart static "${OBJS[@]}" -o "$OUT" -lpthread -ldl
echo "built $(readlink -f $OUT)"
