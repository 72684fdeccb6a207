#!/bin/bash
# Round 2, session 3, end-of-session evidence (1 GPU): full GPU suite, smoke(), the driver's bench command, BASELINE config 2, TPS only,
# reference arm, ncu launch list of the bench command (partitions off: ncu dies on green-context streams), ncu --set full of the
# kernels that changed in this session
set -u
TAG=${1:-r3z}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${TAG}_smoke.txt
timeout -k 10 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 (driver command) rc=$?"
timeout -k 10 600 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
timeout -k 10 600 python bench.py --config c2 --nrow 8192 --ncol 8192 --knots 5000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tpsonly.json 2> gpurun_out/${TAG}_bench_tpsonly.err; echo "bench tps-only rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
python - <<PY
import json
for name in ("c3", "c2", "tpsonly", "reference"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench_%s.json" % name).read().strip().splitlines()[-1])
        print(name, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "parity", d.get("parity") and (d["parity"].get("max_rel_err"), d["parity"].get("lambda_rel_diff")))
        if d.get("roofline"): print("   roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "north star", d["roofline"]["north_star_kernel"]["kernel"][:14], d["roofline"]["north_star_kernel"]["frac"])
        if d.get("kernels"): print("   ", {k: round(v["ms_per_step"], 2) for k, v in list(d["kernels"].items())[:10]})
    except Exception as ex:
        print(name, "no json", ex)
PY
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-tiled > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
cap() {  # name, kernel regex, skip
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/${TAG}_prof_$1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-tiled > gpurun_out/${TAG}_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap trees k_ens_trees 1
cap band_solve k_band_solve 1
cap tri_eig k_tri_eig 1
cap leaf k_leaf_stream 1
ls -la gpurun_out/${TAG}_prof_* | head
