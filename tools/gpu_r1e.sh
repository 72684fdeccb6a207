#!/bin/bash
# r1e: where does the time go?  k_sytrd phase timers + ncu --set full of the leaf / fit kernels.
set -u
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout -k 10 200 python tools/fit_check.py persistent 200 1100 5000 > gpurun_out/${TAG}_fit.log 2>&1; echo "fit_check rc=$?"
cat gpurun_out/${TAG}_fit.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream|k_leaf_prep|k_tri_eig" -c 6 -f -o gpurun_out/${TAG}_tps \
  python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_tps.log 2>&1; echo "ncu tps rc=$?"
tail -3 gpurun_out/${TAG}_ncu_tps.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream|k_sytrd" -c 3 -f -o gpurun_out/${TAG}_c3 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
tail -3 gpurun_out/${TAG}_ncu_c3.log
ls -la gpurun_out/
