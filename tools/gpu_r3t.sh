#!/bin/bash
# Round 2, session 3: TPS-only and config-3 lines with the trimmed bulge chase
set -u
TAG=${1:-r3t}
mkdir -p gpurun_out
timeout -k 10 100 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_tpsonly.json 2> gpurun_out/${TAG}_bench_tpsonly.err; echo "bench tps-only rc=$?"
timeout -k 10 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tiled > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
for name in ("tpsonly", "c3"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench_%s.json" % name).read().strip().splitlines()[-1])
        print(name, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "chase", round(d["kernels"]["k_sbr_chase_ll"]["ms_per_step"], 2), "svm", d["kernels"].get("k_ens_svm_tma", {}).get("ms_per_step"))
    except Exception as ex:
        print(name, "no json", ex)
PY
