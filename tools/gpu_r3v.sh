#!/bin/bash
# Round 2, session 3: k_tri_eig with sign-bit counting - fit tests (incl. the 5 000-knot fit against the oracle), TPS-only and config-3 lines
set -u
TAG=${1:-r3v}
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_tps_gpu.py tests/test_config_scale_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_fit.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_fit.txt
timeout -k 10 300 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_tpsonly.json 2> gpurun_out/${TAG}_bench_tpsonly.err; echo "bench tps-only rc=$?"
timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-tiled > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
for name in ("tpsonly", "c3"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench_%s.json" % name).read().strip().splitlines()[-1])
        print(name, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "parity", d.get("parity") and (d["parity"].get("max_rel_err"), d["parity"].get("lambda_rel_diff")), "tri_eig", round(d["kernels"]["k_tri_eig"]["ms_per_step"], 2))
    except Exception as ex:
        print(name, "no json", ex)
PY
