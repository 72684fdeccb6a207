#!/bin/bash
# Round 2, call s (1 GPU): SM partitions (green contexts) for stage 1 || forest kernel
set -u
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ensemble_gpu.py -m gpu -q -x -k "partitions or full_ensemble" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tiled "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"; tail -2 gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:9]}, "lam", d["fit"]["lambda"])
except Exception as ex:
    print("${name} no json", ex)
PY
}
run off --param gc_split=-1
run gc56
run gc40 --param gc_split=40
run gc72 --param gc_split=72
