#!/bin/bash
# developer A/B builds of the library: tools/build_variant.sh NAME file.cu [-DX=..]...  ->  tools/variants/lib_NAME.so
# (loaded with MOPOE_LIB_PATH; the other objects come from the regular in-tree build)
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/tools/variants
obj=$root/tools/variants/${name}_${src%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $root/mopoe_mimic_b200/csrc/$src -o $obj
objs=""
for o in $root/mopoe_mimic_b200/build/*.o; do
  if [ "$(basename $o)" != "${src%.cu}.o" ]; then objs="$objs $o"; fi
done
nvcc -shared -o $root/tools/variants/lib_$name.so $obj $objs
rm -f $obj
echo $root/tools/variants/lib_$name.so
