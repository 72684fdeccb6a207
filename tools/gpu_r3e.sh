#!/bin/bash
# round 2, session 3: forest kernel - balanced CTA prune (tree blocks round-robin over the warps), two chains per lane in the warp prune,
# explicit shared addresses + branch-free per-cell walk; k_tri_eig as one wave with integer sign logic
set -u
TAG=${1:-r3e}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ensemble_gpu.py tests/test_tps_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_ens_tps.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_ens_tps.txt
timeout -k 10 400 python tools/ens_check.py both --kept rb,bgnmrv --fuse 2 --levels 2 --svm 0 > gpurun_out/${TAG}_ens_check.txt 2>&1; echo "ens_check rc=$?"; cat gpurun_out/${TAG}_ens_check.txt | tail -12
timeout -k 10 600 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tpsonly.json 2> gpurun_out/${TAG}_bench_tpsonly.err; echo "bench tps-only rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_tpsonly.json").read().strip().splitlines()[-1])
print("tps-only value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "parity", d.get("parity") and d["parity"].get("max_rel_err"))
print({k: round(v["ms_per_step"], 2) for k, v in list(d["kernels"].items())[:14]})
PY
