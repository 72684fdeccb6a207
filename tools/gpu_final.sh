#!/bin/bash
# end-of-round evidence: tests, smoke, benches, ncu launch list + full capture of the hot kernels
set -u
TAG=${1:-r1s}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -9 gpurun_out/${TAG}_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
tail -2 gpurun_out/${TAG}_bench_c3.err
timeout -k 10 400 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tps8192.json 2> gpurun_out/${TAG}_bench_tps8192.err; echo "bench tps rc=$?"
timeout -k 10 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream|k_ens_trees|k_ens_svm|k_syrk_dmma|k_leaf_prep" -c 12 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
# two-stage fit: the first panels (largest trailing matrix) of stage 1, then the bulge chase on its own
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_sbr_qr|k_sbr_av|k_sbr_r2k|k_sbr_vtz|k_sbr_w|k_sbr_pu" -c 14 -f -o gpurun_out/${TAG}_prof_sbr \
  python tools/fit_check.py default 5000 > gpurun_out/${TAG}_ncu_sbr.log 2>&1; echo "ncu sbr rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_sbr_chase" -c 1 -f -o gpurun_out/${TAG}_prof_chase \
  python tools/fit_check.py default 5000 > gpurun_out/${TAG}_ncu_chase.log 2>&1; echo "ncu chase rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_stream" -c 2 -f -o gpurun_out/${TAG}_prof_tps \
  python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full_tps.log 2>&1; echo "ncu full tps rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_tps8192.json", "gpurun_out/${TAG}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 2), "parity", d.get("parity"))
        if d.get("roofline"): print("  roofline", d["roofline"]["kernel"][:20], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
        for k, v in list((d.get("kernels") or {}).items())[:10]:
            print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
        print("  cpu", d.get("cpu_baseline"))
    except Exception as e:
        print("no bench json", f, e)
PY
