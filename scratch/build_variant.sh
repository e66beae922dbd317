#!/bin/bash
# usage: scratch/build_variant.sh NAME "-DS3G_SWT=256 ..."   -> scratch/variants/libNAME.so
set -e
cd /root/repo/starch3_b200/csrc
N=$1; shift
O=../../build/obj_$N; mkdir -p $O ../../scratch/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in ${FILES:-bwt}; do
  /usr/local/cuda/bin/nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr "$@" -c $f.cu -o $O/$f.o 2> $O/$f.ptxas.log || (cat $O/$f.ptxas.log; false)
done
B=../../build/obj
/usr/local/cuda/bin/nvcc $ARCH -shared -cudart static -o ../../scratch/variants/lib$N.so $(for f in api tokenize_transform rle_crc bwt bwt_bucket mtf_huff assemble shard multi decode bzshim; do if [ -f $O/$f.o ]; then echo $O/$f.o; else echo $B/$f.o; fi; done)
grep -A2 "${KGREP:-k_sweep}" $O/*.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
