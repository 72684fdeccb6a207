#!/bin/bash
# Round 2, call n (1 GPU): full GPU suite; kernel order with the small-footprint chase
set -u
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tiled "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
print("${name}", "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:6]})
PY
}
run svm_first
run trees_first --param ens_order=1
