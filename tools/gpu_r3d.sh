#!/bin/bash
# round 2, session 3: config 3 with the FP16 split-operand ksvm kernel; ncu --set full of that kernel
set -u
TAG=${1:-r3d}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ensemble_gpu.py -x -q -m gpu > gpurun_out/${TAG}_pytest_ens.txt 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest_ens.txt
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and (round(d["e2e"]["value"], 1), d["e2e"].get("host_ms_per_step")), "parity", d.get("parity") and d["parity"].get("max_rel_err"))
print({k: round(v["ms_per_step"], 2) for k, v in list(d["kernels"].items())[:12]})
PY
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:k_ens_svm_tma" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_svm_f16 \
    python tools/ens_check.py synthetic --kept v --svm 3 --reps 1 > gpurun_out/${TAG}_ncu_svm_f16.log 2>&1; echo "ncu rc=$?"
