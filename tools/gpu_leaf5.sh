#!/bin/bash
set -u
TAG=${1:-leafv}
mkdir -p gpurun_out
timeout -k 10 600 python tools/leaf_check.py --reps 30 --param leaf_impl=2,1 > gpurun_out/${TAG}_leaf_check.txt 2>&1; echo "leaf_check rc=$?"; grep '"k_leaf"' gpurun_out/${TAG}_leaf_check.txt | cut -c1-170
