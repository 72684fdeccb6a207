#!/bin/bash
# round 2, session 3: SM-partition sizes again (the balance between the fit chain and the ensemble chain has moved), forest kernel alone
set -u
TAG=${1:-r3g}
mkdir -p gpurun_out
timeout -k 10 200 python tools/ens_check.py synthetic --kept rb --levels 2 > gpurun_out/${TAG}_ens_check.txt 2>&1; echo "ens_check rc=$?"; tail -n 2 gpurun_out/${TAG}_ens_check.txt
run() {  # name, params...
  local name=$1; shift
  timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-tiled "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    k = d["kernels"]
    print("${name}", "ms", round(d["ms_per_step"], 2), {n: round(k[n]["ms_per_step"], 1) for n in ("k_ens_trees", "k_ens_svm_tma", "k_sbr_chase_ll", "k_sbr_av", "k_sbr_r2k") if n in k})
except Exception as ex:
    print("${name}", "no json", ex)
PY
}
run gc64 --param gc_split=64
run gc80 --param gc_split=80
run gc72_s85 --param gc_share=85
run gc64_s90 --param gc_split=64 --param gc_share=90
run gc56 --param gc_split=56
