#!/bin/bash
# Round 2, call g (1 GPU): TMA start-coordinate probe, TMA-staged ksvm on windows, full GPU suite, ensemble A/B, config-3 bench
set -u
TAG=${1:-r2g}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_tma_probe.txt
for c0 in 0 1 2 3 4 5 8 13 -1 -4 200 223; do for r0 in 0 7 155 -3; do
  timeout 60 tools/bin/tma_probe $c0 $r0 >> gpurun_out/${TAG}_tma_probe.txt 2>&1
done; done
cat gpurun_out/${TAG}_tma_probe.txt
timeout -k 10 300 python tools/dbg_tma.py > gpurun_out/${TAG}_dbg_tma.log 2>&1; echo "dbg_tma rc=$?"; tail -14 gpurun_out/${TAG}_dbg_tma.log
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python tools/ens_check.py both --fuse 2 > gpurun_out/${TAG}_ens_check.log 2>&1; echo "ens_check rc=$?"; cat gpurun_out/${TAG}_ens_check.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["e2e"]["ms_per_step"], "tiled", d["mltps_tiled"]["ms_per_step"])
print("  parity", d.get("parity"))
for k, v in list((d.get("kernels") or {}).items())[:14]:
    print("    ", k, round(v["ms_per_step"], 3), v.get("hbm_frac"))
PY
