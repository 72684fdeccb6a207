#!/bin/bash
# Round 2, call v (1 GPU): grid cap of the bulge chase (SM-time it takes from the ensemble kernels)
set -u
TAG=${1:-r2v}
mkdir -p gpurun_out
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tiled --no-e2e "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"; tail -2 gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms", round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:6]})
except Exception as ex:
    print("${name} no json", ex)
PY
}
run gc72_g64 --param gc_split=72 --param sbr_chase_ctas=64
run gc72_g48 --param gc_split=72 --param sbr_chase_ctas=48
run gc72_g40 --param gc_split=72 --param sbr_chase_ctas=40
run gc72_g32 --param gc_split=72 --param sbr_chase_ctas=32
run gc72_g24 --param gc_split=72 --param sbr_chase_ctas=24
run tps_g48 --config c2 --nrow 8192 --ncol 8192 --knots 5000 --param sbr_chase_ctas=48
run tps_g32 --config c2 --nrow 8192 --ncol 8192 --knots 5000 --param sbr_chase_ctas=32
