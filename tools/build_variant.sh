#!/bin/sh
# Build a tuning variant of the library: tools/build_variant.sh NAME -DPICHA_FAST_G=4 ...
# -> build/variants/libpicha_b200_NAME.so   (use with PICHA_B200_LIB=...)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/variants; mkdir -p $out/$name
cd $root/picha_b200/csrc
for f in $(ls *.cu | sed "s/\.cu$//"); do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off "$@" -c $f.cu -o $out/$name/$f.o &
done
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c tables.cc -o $out/$name/tables.o
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpicha_b200_$name.so $out/$name/*.o
echo $out/libpicha_b200_$name.so
