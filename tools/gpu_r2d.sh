#!/bin/bash
# Round 2, call d (1 GPU): ensemble kernels A/B (two-level forests, fused kernel) on synthetic and real rasters, ncu of the forest and fused kernels
set -u
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ensemble_gpu.py tests/test_config_scale_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout -k 10 900 python tools/ens_check.py both > gpurun_out/${TAG}_ens_check.log 2>&1; echo "ens_check rc=$?"; cat gpurun_out/${TAG}_ens_check.log
for lv in 1 2; do
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_ens_trees" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_trees_l${lv} \
  python tools/ens_check.py synthetic --levels ${lv} --svm 1 --fuse 2 --reps 1 > gpurun_out/${TAG}_ncu_trees_l${lv}.log 2>&1; echo "ncu trees l${lv} rc=$?"
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_ens_fused" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_fused \
  python tools/ens_check.py synthetic --levels 2 --svm 1 --fuse 1 --reps 1 > gpurun_out/${TAG}_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["e2e"]["ms_per_step"], "tiled", d["mltps_tiled"]["ms_per_step"])
print("  parity", d.get("parity"))
for k, v in list((d.get("kernels") or {}).items())[:10]:
    print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
PY
