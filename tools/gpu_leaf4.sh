#!/bin/bash
set -u
TAG=${1:-leafn}
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_warp" -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_leaf_warp \
  python tools/leaf_check.py --reps 2 > gpurun_out/${TAG}_ncu_leaf_warp.log 2>&1; echo "ncu rc=$?"
