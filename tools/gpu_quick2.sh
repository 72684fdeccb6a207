#!/bin/bash
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 10 400 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_tps8192.json 2> gpurun_out/${TAG}_bench_tps8192.err; echo "bench tps rc=$?"
timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_tps8192.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2))
        print("  roofline", d["roofline"]["kernel"][:20], d["roofline"]["frac"])
        for k, v in d["kernels"].items():
            if "leaf" in k or "ens" in k or "sytrd" in k: print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
    except Exception as e:
        print("no bench json", f, e)
PY
