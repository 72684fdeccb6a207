#!/bin/bash
# GCV / Cholesky fit paths with kernel timers (k_sytrd phase timers in persistent mode) + the fit tests
set -u
TAG=${1:-fit}
mkdir -p gpurun_out
L=gpurun_out/${TAG}_fit.log; : > $L
for mode in persistent chol; do
timeout -k 10 200 python tools/fit_check.py $mode 35 200 1100 5000 >> $L 2>&1; echo "fit_check $mode rc=$?" | tee -a $L
done
cat $L
timeout -k 10 300 python -m pytest tests/test_tps_gpu.py -m gpu -x -q -k "fit" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
