#!/bin/bash
# A/B builds of the two-tile star kernel: tools/pp_variants.sh name "-DPP_RING=4 -DPP_LDCS=0" ...  -> csrc/_var/lib_<name>.so
# (the other objects come from the last product build)
set -e
cd "$(dirname "$0")/../deepsc-gan_b200/csrc"
mkdir -p _var
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I. $flags -c -o _var/pp_$name.o dsc_star_pp.cu
  objs=$(ls _obj/*.o | grep -v "dbg_" | grep -v "dsc_star_pp.o")
  nvcc -shared -o _var/lib_$name.so $objs _var/pp_$name.o
  echo built _var/lib_$name.so
done
