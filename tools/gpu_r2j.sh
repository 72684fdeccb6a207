#!/bin/bash
# Round 2, call j (1 GPU): ncu --set full of the grid-evaluation kernel (fused and TPS-only) after the 2-D tensor copy
set -u
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream" -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_leaf_fused \
  python tools/leaf_check.py --reps 2 > gpurun_out/${TAG}_ncu_leaf_fused.log 2>&1; echo "ncu leaf fused rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream" -s 14 -c 1 -f -o gpurun_out/${TAG}_prof_leaf_tps \
  python tools/leaf_check.py --reps 2 > gpurun_out/${TAG}_ncu_leaf_tps.log 2>&1; echo "ncu leaf tps rc=$?"
tail -3 gpurun_out/${TAG}_ncu_leaf_tps.log
