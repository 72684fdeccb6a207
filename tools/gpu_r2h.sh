#!/bin/bash
# Round 2, call h (1 GPU): config-3 bench with the ksvm kernel first (and the other order for comparison), ncu --set full of the grid-evaluation kernel
set -u
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --param ens_order=1 --no-cpu-baseline --no-e2e --no-tiled > gpurun_out/${TAG}_bench_c3_trees_first.json 2> gpurun_out/${TAG}_bench_c3_trees_first.err; echo "bench c3 trees first rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_c3_trees_first.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 2))
    for k, v in list((d.get("kernels") or {}).items())[:8]:
        print("    ", k, round(v["ms_per_step"], 3))
PY
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream" -c 2 -f -o gpurun_out/${TAG}_prof_leaf \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-tiled > gpurun_out/${TAG}_ncu_leaf.log 2>&1; echo "ncu leaf rc=$?"
