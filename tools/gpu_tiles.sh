#!/bin/bash
set -u
TAG=${1:-tiles}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --config c4 --steps 2 --warmup 2 > gpurun_out/${TAG}_c4_n1.json 2> gpurun_out/${TAG}_c4_n1.err; echo "bench c4 N=1 rc=$?"
tail -3 gpurun_out/${TAG}_c4_n1.err
for tr in 1 4; do
timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --tree-rows $tr > gpurun_out/${TAG}_c3_tr${tr}.json 2> gpurun_out/${TAG}_c3_tr${tr}.err; echo "bench c3 tree_rows=$tr rc=$?"
done
python - <<PY
import json
for f in ("gpurun_out/${TAG}_c4_n1.json", "gpurun_out/${TAG}_c3_tr1.json", "gpurun_out/${TAG}_c3_tr4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2))
        k = d.get("kernels") or d.get("kernels_rank0")
        print("   ", {a: round(b["ms_per_step"], 1) for a, b in list(k.items())[:9]})
    except Exception as e:
        print("no json", f, e)
PY
