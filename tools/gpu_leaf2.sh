#!/bin/bash
set -u
TAG=${1:-leafy}
mkdir -p gpurun_out
timeout -k 10 600 python tools/leaf_check.py --reps 30 > gpurun_out/${TAG}_leaf_check.txt 2>&1; echo "leaf_check rc=$?"; cat gpurun_out/${TAG}_leaf_check.txt | tail -4 | cut -c1-200
timeout -k 10 900 python -m pytest tests/test_tps_gpu.py tests/test_ensemble_gpu.py tests/test_tiles_gpu.py tests/test_config_scale_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-tiled > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 2), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "north star", d["roofline"]["north_star_kernel"]["frac"])
PY
