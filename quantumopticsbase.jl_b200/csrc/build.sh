#!/bin/bash
# Build libqob200.so (sm_100a only) in-tree next to the package.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libqob200.so"
NVCC="${NVCC:-nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=default
       --fmad=true -cudart static)
mkdir -p "$HERE/../build"
OBJS=()
for f in qob_api qob_kernels_gather qob_kernels_qtile qob_kernels_qreg qob_kernels_dtile qob_kernels_axis qob_kernels_lindblad qob_kernels_ptrace qob_dist; do
  if [ ! -f "$HERE/../build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/../build/$f.o" ] || [ "$HERE/qob_internal.h" -nt "$HERE/../build/$f.o" ] || [ "$HERE/../../include/qob200.h" -nt "$HERE/../build/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" "$@" -c "$HERE/$f.cu" -o "$HERE/../build/$f.o" &
  fi
  OBJS+=("$HERE/../build/$f.o")
done
wait
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static "${OBJS[@]}" -o "$OUT"
echo "built $OUT"
