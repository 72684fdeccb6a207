#!/bin/bash
# Round 2, session 3: last check of HEAD - ensemble tests (incl. the two-pass forest test), smoke(), the driver's bench command
set -u
TAG=${1:-r3y}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ensemble_gpu.py tests/test_tps_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_ens_tps.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_ens_tps.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${TAG}_smoke.txt
timeout -k 10 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "parity", d["parity"]["max_rel_err"], d["parity"]["lambda_rel_diff"], "clocks", d["clocks"])
print({k: round(v["ms_per_step"], 2) for k, v in list(d["kernels"].items())[:10]})
PY
