#!/bin/bash
# First GPU call of the next round: (1) the driver's own checks, (2) the experimental kernels - tensor-core ksvm (k_ens_svm_mma,
# svm_impl = 1: parity test + bench) and the band-form coefficient solve (coef_impl = 1: parity test, timing in sbr_check.py), (3) the evidence that could not be taken in round 1 (ncu launch list with the two-stage fit).
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
MB_EXPERIMENTAL=1 timeout -k 10 200 python -m pytest tests/test_ensemble_gpu.py -m gpu -q -k svm_tensor > gpurun_out/${TAG}_pytest_svm_mma.log 2>&1; echo "svm_mma pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_svm_mma.log
MB_EXPERIMENTAL=1 timeout -k 10 200 python -m pytest tests/test_tps_gpu.py -m gpu -q -k "band_form or watcher or fused_panel" > gpurun_out/${TAG}_pytest_band_form.log 2>&1; echo "band_form pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_band_form.log
timeout -k 10 200 python tools/sbr_check.py 1100 5000 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; cat gpurun_out/${TAG}_check.log
timeout -k 10 300 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
timeout -k 10 300 python bench.py --param svm_impl=1 > gpurun_out/${TAG}_bench_c3_svm_mma.json 2> gpurun_out/${TAG}_bench_c3_svm_mma.err; echo "bench svm_mma rc=$?"
timeout -k 10 300 python bench.py --param svm_impl=2 > gpurun_out/${TAG}_bench_c3_svm_mma_poly.json 2> gpurun_out/${TAG}_bench_c3_svm_mma_poly.err; echo "bench svm_mma_poly rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import json
for f in ("bench_c3", "bench_c3_svm_mma", "bench_c3_svm_mma_poly"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "parity", d.get("parity"))
        print("    ", {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:12]})
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
