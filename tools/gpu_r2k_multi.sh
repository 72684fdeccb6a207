#!/bin/bash
# Round 2, multi-GPU call k (gpurun --gpus N): sharded tiles.merge (tests), config 4 with the seam-strip exchange, config 5 by layers,
# the default arm (one raster in N row blocks, weak) and the reference arm
set -u
N=${1:-2}
TAG=${2:-r2k}
WHAT=${3:-all}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$WHAT" = "all" ]; then
timeout -k 10 900 python -m pytest tests/test_comm_gpu.py tests/test_tiles_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_comm_tiles.log 2>&1; echo "pytest comm+tiles rc=$?"; tail -5 gpurun_out/${TAG}_pytest_comm_tiles.log
fi
run() {  # name, extra args
  local name=$1; shift
  timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" \
    > gpurun_out/${TAG}_bench_${name}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${name}_${N}gpu.err; echo "bench $name rc=$?"; tail -3 gpurun_out/${TAG}_bench_${name}_${N}gpu.err
}
run c4 --config c4 --steps 3 --warmup 2
run c5 --config c5 --steps 2 --warmup 1
run weak --steps 5 --warmup 3
if [ "$WHAT" = "all" ]; then
run reference --impl reference --steps 1 --warmup 1
fi
python - <<PY
import json
for name in ("c4", "c5", "weak", "reference"):
    f = "gpurun_out/${TAG}_bench_%s_${N}gpu.json" % name
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(name, "n_gpus", d["n_gpus"], "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "scaling", d["scaling"], "e2e", d.get("e2e") and round(d["e2e"]["value"], 2))
        print("   parity", d.get("parity")); print("   par", d["config"].get("parallelism")); print("   coll", d.get("collectives"))
    except Exception as e:
        print("no bench json", f, e)
PY
