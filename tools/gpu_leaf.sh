#!/bin/bash
set -u
TAG=${1:-leafx}
mkdir -p gpurun_out
timeout -k 10 600 python tools/leaf_check.py --reps 30 > gpurun_out/${TAG}_leaf_check.txt 2>&1; echo "leaf_check rc=$?"; cat gpurun_out/${TAG}_leaf_check.txt | tail -4 | cut -c1-200
