#!/bin/bash
# Round 2, session 3: final code on 2 GPUs - the GPU suite including the 2-rank tests, then the default arm (one raster in two row blocks), weak and strong
set -u
N=${1:-2}
TAG=${2:-r3w}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_${N}gpu.txt 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_${N}gpu.txt
for sc in weak strong; do
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --scaling $sc \
  > gpurun_out/${TAG}_bench_${sc}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${sc}_${N}gpu.err; echo "bench $sc rc=$?"; tail -n 2 gpurun_out/${TAG}_bench_${sc}_${N}gpu.err
done
python - <<PY
import json
for sc in ("weak", "strong"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench_%s_${N}gpu.json" % sc).read().strip().splitlines()[-1])
        print(sc, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "parity", d.get("parity") and d["parity"]["max_rel_err"])
        print("   ", {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:6]})
    except Exception as ex:
        print(sc, "no json", ex)
PY
