#!/bin/bash
# A/B builds of the experimental two-tile star kernel (csrc/debug/dsc_star_pp.cu, debug-tools library):
#   tools/pp_variants.sh name "-DPP_RING=4" ...  ->  csrc/_var/lib_<name>.so  (bind with DSC_LIB_PATH=...; tools/pp_check.py)
# The other objects come from the last debug build (python deepsc-gan_b200/build.py --debug).
set -e
cd "$(dirname "$0")/../deepsc-gan_b200/csrc"
mkdir -p _var
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I. -DDSC_DEBUG_TOOLS=1 $flags -c -o _var/pp_$name.o debug/dsc_star_pp.cu
  objs=$(ls _obj/dbg_*.o | grep -v "dbg_dsc_star_pp.o")
  nvcc -shared -o _var/lib_$name.so $objs _var/pp_$name.o 2>/dev/null
  echo built _var/lib_$name.so
done
