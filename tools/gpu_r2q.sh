#!/bin/bash
# Round 2, call q (1 GPU): CTA-wide band solve, fused panel kernels, deferred ensemble join: tests, fit timings, config 3, TPS only
set -u
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_tps_gpu.py tests/test_rshim_gpu.py tests/test_ensemble_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python tools/chase_check.py --sizes 1100,5000 --impls 3 > gpurun_out/${TAG}_fit_check.txt 2>&1; cat gpurun_out/${TAG}_fit_check.txt | tail -3
run() {
  local name=$1; shift
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tiled "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "bench $name rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
print("${name}", "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:12]}, "lam", d["fit"]["lambda"])
PY
}
run c3
run c3_unfused --param sbr_fuse=2
run tpsonly --config c2 --nrow 8192 --ncol 8192 --knots 5000
