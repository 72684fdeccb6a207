#!/bin/bash
# last check of the round: what the driver runs (GPU tests, smoke, default bench), nothing else
set -u
TAG=${1:-last}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout -k 10 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "parity", d["parity"]["max_rel_err"], "roofline", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], "clocks", d["clocks"])
for k, v in list(d["kernels"].items())[:9]: print("    ", k, round(v["ms_per_step"], 2))
print("cpu", d["cpu_baseline"])
PY
