#!/bin/bash
# Round 2, call t (1 GPU): SM partitions x chase variant sweep
set -u
TAG=${1:-r2t}
mkdir -p gpurun_out
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
run gc64 --param gc_split=64
run gc64_ll --param gc_split=64 --param sbr_chase_impl=3
run gc72_ll --param gc_split=72 --param sbr_chase_impl=3
run gc56_ll --param gc_split=56 --param sbr_chase_impl=3
run off_ll --param gc_split=-1 --param sbr_chase_impl=3
run gc64_dec --param gc_split=64 --param sbr_chase_impl=1
