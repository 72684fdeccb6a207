#!/bin/bash
set -u
TAG=${1:-r1f}
mkdir -p gpurun_out
L=gpurun_out/${TAG}_fit.log; : > $L
for mode in persistent phases chol; do
  timeout -k 10 240 python tools/fit_check.py $mode 35 200 1100 5000 >> $L 2>&1; echo "fit_check $mode rc=$?" | tee -a $L
done
cat $L
timeout -k 10 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
tail -3 gpurun_out/${TAG}_bench_c3.err
timeout -k 10 600 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tps8192.json 2> gpurun_out/${TAG}_bench_tps8192.err; echo "bench tps rc=$?"
tail -3 gpurun_out/${TAG}_bench_tps8192.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_tps8192.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), "parity", d["parity"])
        print("  roofline", d["roofline"]["kernel"][:20], d["roofline"]["frac"])
        for k, v in d["kernels"].items():
            print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
    except Exception as e:
        print("no bench json", f, e)
PY
