#!/bin/bash
# Round 2, call c (1 GPU): tests + bench after the two-level forest kernel and the promoted defaults
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
tail -3 gpurun_out/${TAG}_bench_c3.err
timeout -k 10 300 python tools/sbr_check.py 5000 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; cat gpurun_out/${TAG}_check.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 2), d["e2e"]["ms_per_step"])
        print("  parity", d.get("parity"))
        print("  tiled", d.get("mltps_tiled"))
        r = d["roofline"]; print("  roofline", r["kernel"], r["frac"], "| north-star", r["north_star_kernel"]["frac"], "| step", r["step"]["frac"])
        for k, v in list((d.get("kernels") or {}).items())[:14]:
            print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
        print("  cpu", d.get("cpu_baseline"))
    except Exception as e:
        print("no bench json", f, e)
PY
