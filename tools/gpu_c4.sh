#!/bin/bash
set -u
N=${1:-2}; TAG=${2:-c4}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
timeout -k 10 600 python -m pytest tests/test_tiles_gpu.py tests/test_ensemble_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
fi
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --config c4 --steps 2 --warmup 2 > gpurun_out/${TAG}_c4_n${N}.json 2> gpurun_out/${TAG}_c4_n${N}.err; echo "bench c4 N=$N rc=$?"
grep -i "error" gpurun_out/${TAG}_c4_n${N}.err | head -5
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_c4_n${N}.json").read().strip().splitlines()[-1])
    print("c4 N=$N value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "na", d["na_fraction"])
    print("   ", {a: round(b["ms_per_step"], 1) for a, b in list(d["kernels_rank0"].items())[:8]})
except Exception as e:
    print("no json", e)
PY
