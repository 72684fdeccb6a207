#!/bin/bash
# multi-GPU check: default bench (config 3, weak scaling, one 8192^2 tile per GPU) and config 4 (16384^2 tiled across the GPUs)
set -u
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
PORT=29517
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_c3_n${N}.json 2> gpurun_out/${TAG}_c3_n${N}.err; echo "bench c3 N=$N rc=$?"
tail -3 gpurun_out/${TAG}_c3_n${N}.err
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
  bench.py --gpus $N --config c4 --steps 2 --warmup 2 > gpurun_out/${TAG}_c4_n${N}.json 2> gpurun_out/${TAG}_c4_n${N}.err; echo "bench c4 N=$N rc=$?"
tail -5 gpurun_out/${TAG}_c4_n${N}.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_c3_n${N}.json", "gpurun_out/${TAG}_c4_n${N}.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "n_gpus", d["n_gpus"], "scaling", d["scaling"], "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"],1))
        k = d.get("kernels") or d.get("kernels_rank0")
        print("   ", {a: round(b["ms_per_step"], 1) for a, b in list(k.items())[:8]})
    except Exception as e:
        print("no json", f, e)
PY
