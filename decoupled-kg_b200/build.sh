#!/usr/bin/env bash
# Build libdkg_b200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/csrc"
OUT="$HERE/lib"
mkdir -p "$OUT" "$OUT/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr ${DKG_NVCC_EXTRA:-})
pids=()
for f in dkg_api dkg_prepare dkg_gemm dkg_forward dkg_emax dkg_coupled dkg_ozaki; do
  if [[ ! -f "$OUT/obj/$f.o" || "$SRC/$f.cu" -nt "$OUT/obj/$f.o" || -n "$(find "$SRC" "$HERE/../include" -name '*.cuh' -newer "$OUT/obj/$f.o" -o -name '*.h' -newer "$OUT/obj/$f.o")" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$SRC/$f.cu" -o "$OUT/obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libdkg_b200.so" "$OUT"/obj/*.o -lcudart
echo "built $OUT/libdkg_b200.so"
