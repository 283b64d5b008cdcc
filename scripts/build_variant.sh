#!/bin/bash
# Development tool: build an experimental variant of the library with extra -D flags next to the product library, e.g.
#   scripts/build_variant.sh u4 -DNRC_INFER_UNROLL=4
# -> nrc_hpm_renderer_b200/variants/libnrchpm_b200_u4.so; select it at run time with NRCHPM_LIB=<path> (see _lib.py).
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
src=$here/nrc_hpm_renderer_b200/csrc
out=$here/nrc_hpm_renderer_b200/variants
mkdir -p "$out" /tmp/nrchpm_variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH "$@" -c "$src/nrc.cu" -o /tmp/nrchpm_variants/nrc_$name.o
[ -f "$src/hpm.o" ] || make -C "$src" hpm.o
nvcc -shared $ARCH -o "$out/libnrchpm_b200_$name.so" /tmp/nrchpm_variants/nrc_$name.o "$src/hpm.o"
echo "$out/libnrchpm_b200_$name.so"
