#!/bin/bash
# Round 2, call u (1 GPU): SM partitions: rows of the forest kernel on the ensemble partition, chase variants
set -u
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ensemble_gpu.py tests/test_tps_gpu.py -m gpu -q -x -k "partitions or full_ensemble or chase_variants" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tiled --no-e2e "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"; tail -2 gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms", round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:8]})
except Exception as ex:
    print("${name} no json", ex)
PY
}
run gc72_s78 --param gc_split=72
run gc72_s65 --param gc_split=72 --param gc_share=65
run gc72_s90 --param gc_split=72 --param gc_share=90
run gc64_s78 --param gc_split=64
run gc64_s90 --param gc_split=64 --param gc_share=90
run gc72_s78_ll2 --param gc_split=72 --param sbr_chase_impl=4
run gc72_s78_c2 --param gc_split=72 --param sbr_chase_impl=2
run gc80_s78 --param gc_split=80
