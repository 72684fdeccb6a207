#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, bench, ncu launch list, ncu full capture of the hot kernels.
# Usage: tools/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KREG=${2:-'k_leaf|k_ens_trees|k_ens_svm|k_ens_final'}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_tps8192.json 2> gpurun_out/${TAG}_bench_tps8192.err; echo "bench tps rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${KREG}" -c 8 -o gpurun_out/${TAG}_prof -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/${TAG}_bench_c3.json | head -c 3000
