#!/bin/bash
# round 2, session 3: forest kernel tuning constants (A/B copies of the library, tools/build_variants.sh), synthetic and real rasters
set -u
TAG=${1:-r3h}
mkdir -p gpurun_out
for v in base ilp6 ilp8c4 ilp8c3 w3 p4c4; do
  if [ $v = base ]; then unset MB_LIB; else export MB_LIB=$PWD/machisplin_b200/build/variants/lib_$v.so; fi
  timeout -k 10 200 python tools/ens_check.py both --kept rb --levels 2 > gpurun_out/${TAG}_ens_check_$v.txt 2>&1; echo "$v rc=$?"; grep "kept=rb" gpurun_out/${TAG}_ens_check_$v.txt | cut -c1-20,95-140
done
