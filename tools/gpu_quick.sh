#!/bin/bash
# quick GPU check: tests + short benches (no ncu).  Usage: tools/gpu_quick.sh <tag> [extra bench args]
set -u
TAG=${1:-q}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 "$@" > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
tail -5 gpurun_out/${TAG}_bench_c3.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_c3.json"))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"],"parity",d["parity"])
    print("roofline",d["roofline"])
    for k,v in d["kernels"].items(): print("  ",k,round(v["ms_per_step"],3),v.get("hbm_frac"))
    print("cpu",d["cpu_baseline"])
except Exception as e: print("no bench json",e)
PY
