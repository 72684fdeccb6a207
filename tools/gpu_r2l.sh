#!/bin/bash
# Round 2, call l (1 GPU): what slows the ensemble kernel that runs beside the post-stage-1 fit?  A/B of the chase variants
set -u
TAG=${1:-r2l}
mkdir -p gpurun_out
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-tiled "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
print("${name}", "ms", round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:9]})
PY
}
run default
run chase2 --param sbr_chase_impl=2
run coef2 --param coef_impl=2
run fixedlam --lam 0.004
run tpsonly --config c2 --nrow 8192 --ncol 8192 --knots 5000
