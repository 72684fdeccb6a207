#!/bin/bash
# round 2, session 3: ksvm kernel - FP16 split operands (2 HMMA) against 3 x TF32 (3 HMMA), exponentials shared between MUFU.EX2 and the FMA pipe (svm_poly)
set -u
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ensemble_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_ens.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_ens.txt
timeout -k 10 400 python tools/ens_check.py synthetic --kept v,bgnmrv --fuse 2 --levels 2 --svm ${2:-1,3} --poly ${3:-1,2,3} > gpurun_out/${TAG}_ens_check.txt 2>&1; echo "ens_check rc=$?"; cat gpurun_out/${TAG}_ens_check.txt | tail -30
