#!/bin/bash
# Builds style_transfer_b200/libstyle_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libstyle_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -cudart static -I"$HERE/../../include")
mkdir -p "$HERE/build"
pids=()
for f in engine kernels_simt kernels_image conv_tc conv_tc2 gram_tc conv_first_tc conv_pix_tc; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || \
     [ -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$HERE/build/$f.o" 2>/dev/null)" ]; then
    "$NVCC" "${FLAGS[@]}" ${EXTRA_NVCC_FLAGS:-} -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -cudart static -o "$OUT" "$HERE"/build/{engine,kernels_simt,kernels_image,conv_tc,conv_tc2,gram_tc,conv_first_tc,conv_pix_tc}.o
echo "built $OUT"
