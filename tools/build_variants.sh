#!/bin/bash
# A/B copies of the library: ensemble.cu recompiled with other tuning constants, linked with the objects of the last regular build.
#   tools/build_variants.sh name "-DMB_TREE_ILP=8 -DMB_TREE_CTAS=4" [name2 "flags2" ...]   ->  machisplin_b200/build/variants/lib_<name>.so
# Run one with MB_LIB=machisplin_b200/build/variants/lib_<name>.so python tools/ens_check.py ...
set -e
cd "$(dirname "$0")/.."
python -m machisplin_b200.build > /dev/null
B=machisplin_b200/build
mkdir -p $B/variants
CUDA_LIB=$(dirname $(dirname $(readlink -f $(which nvcc))))/lib64
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr -Xptxas -v $flags \
    -c machisplin_b200/csrc/ensemble.cu -o $B/variants/ensemble_$name.o 2> $B/variants/ptxas_$name.log
  objs=$(ls $B/*.cu.o | grep -v ensemble.cu.o)
  nvcc -shared -o $B/variants/lib_$name.so $objs $B/variants/ensemble_$name.o -L$CUDA_LIB -lcudart -ldl -Xlinker -rpath,$CUDA_LIB
  echo "$name: $(grep -A2 'k_ens_treesILi1' $B/variants/ptxas_$name.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'k_ens_treesILi1' $B/variants/ptxas_$name.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
