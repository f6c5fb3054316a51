#!/bin/sh
# Quick tuning variant of the downscaling kernel only: tools/build_down_variant.sh NAME -DPICHA_DOWN_NS=3 ...
# Recompiles the launch planner and the u8 rgba / rgb instantiations (cfg3 / cfg5) with the flags and links them
# with the default build's objects (build/csrc, run `make -C picha_b200/csrc` first).
# -> build/variants/libpicha_b200_NAME.so   (use with PICHA_B200_LIB=...)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/variants; mkdir -p $out/$name
cd $root/picha_b200/csrc
units="resize_fast resize_down_u8_c4 resize_down_u8_c3 ${EXTRA_UNITS}"
for f in $units; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v "$@" -c $f.cu -o $out/$name/$f.o 2> $out/$name/$f.log &
done
wait
objs=""
for o in $root/build/csrc/*.o; do
  b=$(basename $o .o)
  case " $units " in *" $b "*) objs="$objs $out/$name/$b.o";; *) objs="$objs $o";; esac
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpicha_b200_$name.so $objs -ldl
rm -f $out/$name/*.o
echo $out/libpicha_b200_$name.so
