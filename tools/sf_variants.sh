#!/bin/bash
# A/B builds of the one-tile star kernel: tools/sf_variants.sh name "-DSF_RING=5" ... -> csrc/_var/lib_<name>.so
set -e
cd "$(dirname "$0")/../deepsc-gan_b200/csrc"
mkdir -p _var
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I. $flags -c -o _var/sf_$name.o dsc_star_fused.cu
  objs=$(ls _obj/*.o | grep -v "dbg_" | grep -v "/dsc_star_fused.o")
  nvcc -shared -o _var/lib_$name.so $objs _var/sf_$name.o 2>/dev/null
  echo built _var/lib_$name.so
done
