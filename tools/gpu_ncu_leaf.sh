#!/bin/bash
set -u
TAG=${1:-leaf}
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream" -c 2 -f -o gpurun_out/${TAG}_tps \
  python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_tps.log 2>&1; echo "ncu tps rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream" -c 2 -f -o gpurun_out/${TAG}_c3 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
