#!/bin/bash
# Round 2, call b (1 GPU): every GPU test incl. the config-scale parity tests and the R shim, the new bench line, and the ncu
# --set full captures of the fit kernels that had only event timers so far (two-stage tridiagonalisation, Cholesky).
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${TAG}_pytest.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
tail -3 gpurun_out/${TAG}_bench_c3.err
# fit kernels: stage 1 around panel 8 (large trailing matrix), the bulge chase, the Cholesky at about a third of the way
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_sbr_qr|k_sbr_av|k_sbr_r2k|k_sbr_vtz|k_sbr_st|k_sbr_w|k_sbr_pu" -s 56 -c 8 -f -o gpurun_out/${TAG}_prof_sbr \
  python tools/fit_check.py default 5000 > gpurun_out/${TAG}_ncu_sbr.log 2>&1; echo "ncu sbr rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_sbr_chase" -c 1 -f -o gpurun_out/${TAG}_prof_chase \
  python tools/fit_check.py default 5000 > gpurun_out/${TAG}_ncu_chase.log 2>&1; echo "ncu chase rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:k_potrf_diag|k_trsm_panel|k_syrk_dmma|k_chol_sweep|k_tri_eig" -s 80 -c 9 -f -o gpurun_out/${TAG}_prof_chol \
  python tools/fit_check.py default 5000 > gpurun_out/${TAG}_ncu_chol.log 2>&1; echo "ncu chol rc=$?"
ls -la gpurun_out/${TAG}_prof*.ncu-rep
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 2))
        print("  parity", d.get("parity"))
        print("  tiled", d.get("mltps_tiled"))
        r = d["roofline"]; print("  roofline", r["kernel"], r["frac"], "| north-star", r["north_star_kernel"]["frac"], "| step", r["step"]["frac"])
        for k, v in list((d.get("kernels") or {}).items())[:12]:
            print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
        print("  cpu", d.get("cpu_baseline"))
    except Exception as e:
        print("no bench json", f, e)
PY
