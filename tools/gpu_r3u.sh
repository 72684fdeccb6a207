#!/bin/bash
# Round 2, session 3: tagged-element bulge chase with trimmed data movement against the flag kernel - bit-identity at every size and kernel time
set -u
TAG=${1:-r3u}
mkdir -p gpurun_out
timeout -k 10 200 python tools/chase_check.py --sizes 36,70,200,1100,5000 --impls 2,3 --reps 3 > gpurun_out/${TAG}_chase_check.txt 2>&1; echo "chase_check rc=$?"; cat gpurun_out/${TAG}_chase_check.txt | cut -c1-200
