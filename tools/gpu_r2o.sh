#!/bin/bash
set -u
TAG=${1:-r2o}
mkdir -p gpurun_out
timeout -k 10 900 python tools/chase_check.py > gpurun_out/${TAG}_chase_check.txt 2>&1; echo "chase_check rc=$?"; cat gpurun_out/${TAG}_chase_check.txt | tail -20
