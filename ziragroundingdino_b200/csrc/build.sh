#!/bin/bash
# Builds ziragroundingdino_b200/_lib/libmsda_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../_lib"
mkdir -p "$OUT"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr ${MSDA_NVCC_EXTRA}"
pids=()
for f in msda_core msda_scatter_mma msda_scatter_mma2 capi probe $(ls "$HERE" | grep -E '^(proj_|zira_|layer_).*\.cu$' | sed 's/\.cu$//'); do
  $NVCC $FLAGS -c "$HERE/$f.cu" -o "$OUT/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -Wno-deprecated-gpu-targets -shared -o "$OUT/libmsda_b200.so" "$OUT"/*.o -lcudart -lcuda
echo "built $OUT/libmsda_b200.so"
