#!/bin/bash
# two-stage fit as the default: GPU tests, smoke, two-stage check at 1100 / 5000 knots, config-3 and TPS-only benches
set -u
TAG=${1:-sbr}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout -k 10 200 python tools/sbr_check.py 1100 5000 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; cat gpurun_out/${TAG}_check.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout -k 10 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
timeout -k 10 300 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --no-cpu-baseline > gpurun_out/${TAG}_bench_tps8192.json 2> gpurun_out/${TAG}_bench_tps8192.err; echo "bench tps rc=$?"
python - <<PY
import json
for f in ("bench_c3", "bench_tps8192"):
    d = json.loads(open("gpurun_out/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None, "parity", d["parity"]["max_rel_err"], "roofline", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], "clocks", d["clocks"])
    for k, v in list(d["kernels"].items())[:12]: print("    ", k, round(v["ms_per_step"], 2))
PY
