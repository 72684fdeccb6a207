#!/bin/bash
set -u
TAG=${1:-b3}
mkdir -p gpurun_out
for v in "sytrd_mode=0" "sytrd_mode=2" "eigen_impl=1" "sytrd_ctas_per_sm=1"; do
  n=$(echo $v | tr '=' '_')
  timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --param $v > gpurun_out/${TAG}_c3_${n}.json 2> gpurun_out/${TAG}_c3_${n}.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_c3_${n}.json"))
print("$v", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), "launches", d["gpu_launches"])
print("   ", {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:8]})
PY
done
