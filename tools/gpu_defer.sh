#!/bin/bash
# config-3 bench with the ensemble launch deferred behind stage 1 of the fit (default) and not deferred
set -u
TAG=${1:-defer}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ensemble_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
for v in 1 0; do
timeout -k 10 300 python bench.py --no-cpu-baseline --param defer_ensemble=$v > gpurun_out/${TAG}_bench_c3_defer$v.json 2> gpurun_out/${TAG}_bench_c3_defer$v.err; echo "bench defer=$v rc=$?"
done
timeout -k 10 300 python bench.py --no-cpu-baseline --param defer_ensemble=0 --param sbr_qr_grid=1 > gpurun_out/${TAG}_bench_c3_defer0_grid.json 2> gpurun_out/${TAG}_bench_c3_defer0_grid.err; echo "bench grid rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable", ex); continue
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "parity", d["parity"]["max_rel_err"])
    print("    ", {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:10]})
PY
