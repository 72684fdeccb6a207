#!/bin/bash
set -u
TAG=${1:-fa}
mkdir -p gpurun_out
for tr in 1 2; do
timeout -k 10 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tree-rows $tr > gpurun_out/${TAG}_c3_tr${tr}.json 2> gpurun_out/${TAG}_c3_tr${tr}.err; echo "bench c3 tree_rows=$tr rc=$?"
done
timeout -k 10 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tree-rows 1 --param sytrd_mode=2 > gpurun_out/${TAG}_c3_tr1_m2.json 2> gpurun_out/${TAG}_c3_tr1_m2.err; echo "bench c3 tr1 mode2 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_c3_tr1.json", "gpurun_out/${TAG}_c3_tr2.json", "gpurun_out/${TAG}_c3_tr1_m2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"],1))
        k = d.get("kernels") or d.get("kernels_rank0")
        print("   ", {a: round(b["ms_per_step"], 1) for a, b in list(k.items())[:6]})
    except Exception as e:
        print("no json", f, e)
PY
