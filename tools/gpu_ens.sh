#!/bin/bash
set -u
TAG=${1:-ensx}
mkdir -p gpurun_out
timeout -k 10 600 python tools/ens_check.py synthetic --fuse 2 --levels 2 --svm 1 > gpurun_out/${TAG}_ens_check.txt 2>&1; echo "ens_check rc=$?"; cat gpurun_out/${TAG}_ens_check.txt | tail -5
