#!/usr/bin/env bash
# Builds lib/libgpd_b200.so for sm_100a (B200). nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $ARCH ${GPD_NVCC_EXTRA:-}"
"$NVCC" $COMMON -c "$HERE/gpd_f32.cu" -o "$HERE/obj/gpd_f32.o" 2> "$HERE/obj/ptxas_f32.log" &
"$NVCC" $COMMON -fmad=false -c "$HERE/gpd_f64.cu" -o "$HERE/obj/gpd_f64.o" 2> "$HERE/obj/ptxas_f64.log" &
"$NVCC" $COMMON -c "$HERE/gpd_api.cu" -o "$HERE/obj/gpd_api.o" 2> "$HERE/obj/ptxas_api.log" &
fail=0
for j in $(jobs -p); do wait "$j" || fail=1; done
if [ "$fail" != 0 ]; then cat "$HERE"/obj/ptxas_*.log | grep -v "^ptxas info" >&2 || true; exit 1; fi
"$NVCC" -shared $ARCH -o "$OUT/libgpd_b200.so" "$HERE/obj/gpd_f32.o" "$HERE/obj/gpd_f64.o" "$HERE/obj/gpd_api.o" -cudart static
echo "built $OUT/libgpd_b200.so"
