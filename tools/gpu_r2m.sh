#!/bin/bash
# Round 2, call m (1 GPU): back-off in the spin loops of the bulge chase (A/B)
set -u
TAG=${1:-r2m}
mkdir -p gpurun_out
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-tiled "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
print("${name}", "ms", round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:6]}, "lam", d["fit"]["lambda"])
PY
}
run sleep0
run sleep20 --param sbr_chase_sleep=20
run sleep50 --param sbr_chase_sleep=50
run sleep100 --param sbr_chase_sleep=100
run sleep200 --param sbr_chase_sleep=200
run sleep500 --param sbr_chase_sleep=500
run c2sleep100 --param sbr_chase_sleep=100 --param sbr_chase_impl=2
run tpsonly_sleep100 --param sbr_chase_sleep=100 --config c2 --nrow 8192 --ncol 8192 --knots 5000
