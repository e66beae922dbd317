#!/bin/bash
# usage: scratch/build_variant.sh NAME "-DS3G_SWT=256 ..."   -> scratch/variants/libNAME.so
set -e
cd /root/repo/starch3_b200/csrc
N=$1; shift
O=../../build/obj_$N; mkdir -p $O ../../scratch/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in bwt; do
  /usr/local/cuda/bin/nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr "$@" -c $f.cu -o $O/$f.o 2> $O/$f.ptxas.log || (cat $O/$f.ptxas.log; false)
done
B=../../build/obj
/usr/local/cuda/bin/nvcc $ARCH -shared -cudart static -o ../../scratch/variants/lib$N.so $B/api.o $B/tokenize_transform.o $B/rle_crc.o $O/bwt.o $B/mtf_huff.o $B/assemble.o
grep -A2 "k_sweep" $O/bwt.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
