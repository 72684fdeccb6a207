#!/bin/bash
# Round 2, call r (1 GPU): full GPU suite (order as the driver runs it), config 3
set -u
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 2), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "tiled", round(d["mltps_tiled"]["ms_per_step"], 1))
print(" roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "north star", d["roofline"]["north_star_kernel"]["frac"], "step", d["roofline"]["step"]["frac"])
print(" parity", d["parity"]["max_rel_err"], d["parity"]["lambda_rel_diff"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
