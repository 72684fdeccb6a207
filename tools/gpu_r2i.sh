#!/bin/bash
# Round 2, call i (1 GPU): leaf kernel with the 2-D tensor copy (A/B), chase with relaxed polls: tests, fit check, config-3 bench
set -u
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout -k 10 600 python tools/leaf_check.py --param leaf_tma=2,1 > gpurun_out/${TAG}_leaf_check.txt 2>&1; echo "leaf_check rc=$?"; cat gpurun_out/${TAG}_leaf_check.txt | tail -8
timeout -k 10 1500 python -m pytest tests/test_tps_gpu.py tests/test_ensemble_gpu.py tests/test_tiles_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --param ens_order=1 --no-cpu-baseline --no-e2e --no-tiled > gpurun_out/${TAG}_bench_c3_trees_first.json 2> gpurun_out/${TAG}_bench_c3_trees_first.err; echo "bench c3 trees first rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_c3_trees_first.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 2), "parity", d.get("parity") and d["parity"]["max_rel_err"])
    print("   north star", d["roofline"].get("north_star_kernel", {}).get("frac"))
    for k, v in list((d.get("kernels") or {}).items())[:8]:
        print("    ", k, round(v["ms_per_step"], 3))
PY
