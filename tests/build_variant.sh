#!/bin/bash
# build_variant.sh NAME FILE.cu "-DFLAG=.. ..."  -> mtgvision_b200/csrc/variants/libmtgv_NAME.so
# (same objects as libmtgv.so except FILE.cu recompiled with the extra flags; select with MTGV_LIB=...)
set -e
cd "$(dirname "$0")/../mtgvision_b200/csrc"
mkdir -p variants
name=$1; file=$2; flags=$3
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off $flags -c -o variants/${file%.cu}_$name.o $file
objs=""
for f in *.cu; do
  if [ "$f" = "$file" ]; then objs="$objs variants/${file%.cu}_$name.o"; else objs="$objs build/${f%.cu}.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libmtgv_$name.so $objs
echo built variants/libmtgv_$name.so
