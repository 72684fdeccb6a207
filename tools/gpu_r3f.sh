#!/bin/bash
# round 2, session 3: full GPU suite, TPS-only and config-3 lines after the forest-kernel, k_tri_eig and k_band_solve changes
set -u
TAG=${1:-r3f}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest.txt
timeout -k 10 600 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tpsonly.json 2> gpurun_out/${TAG}_bench_tpsonly.err; echo "bench tps-only rc=$?"
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
for name in ("tpsonly", "c3"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench_%s.json" % name).read().strip().splitlines()[-1])
        print(name, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and (round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2)), "parity", d.get("parity") and (d["parity"].get("max_rel_err"), d["parity"].get("lambda_rel_diff")))
        print("  ", {k: round(v["ms_per_step"], 2) for k, v in list(d["kernels"].items())[:14]})
    except Exception as ex:
        print(name, "no json", ex)
PY
