#!/bin/bash
# Build a kernel variant of libntrtsim_b200.so (+ host lib) into variants/<name>/ with extra nvcc flags, for same-box A/B
# timing:  tools/build_variant.sh cu1 -DTGB_CABLE_UNROLL=1 ; TGB_LIB_DIR=$PWD/variants/cu1 python bench.py
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/variants/$name
mkdir -p $out
make -C $root -j8 BUILD=$out/obj LIBDIR=$out TGB_EXTRA="$*" $out/libntrtsim_b200.so $out/libtgb_host.so > $out/build.log 2>&1 || { tail -20 $out/build.log; exit 1; }
rm -rf $out/obj
echo built $out
