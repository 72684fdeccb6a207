#!/bin/bash
set -u
TAG=${1:-t}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --config c5 --steps 1 --warmup 1 > gpurun_out/${TAG}_c5_n1.json 2> gpurun_out/${TAG}_c5_n1.err; echo "bench c5 rc=$?"
tail -3 gpurun_out/${TAG}_c5_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_c5_n1.json").read().strip().splitlines()[-1])
    print("c5 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "lam", [round(x, 5) for x in d["lambda_rank0"][:4]])
    print("   ", {a: round(b["ms_per_step"], 1) for a, b in list(d["kernels_rank0"].items())[:9]})
except Exception as e:
    print("no json", e)
PY
