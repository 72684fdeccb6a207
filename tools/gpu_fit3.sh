#!/bin/bash
set -u
TAG=${1:-fit}
mkdir -p gpurun_out
L=gpurun_out/${TAG}_fit.log; : > $L
for mode in persistent phases; do
timeout -k 10 200 python tools/fit_check.py $mode 35 200 1100 5000 >> $L 2>&1; echo "fit_check $mode rc=$?" | tee -a $L
done
grep -v "^   kernels" $L
timeout -k 10 300 python -m pytest tests/test_tps_gpu.py -m gpu -x -q -k "fit" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
