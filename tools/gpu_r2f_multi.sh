#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the library's own NCCL communicator, two ranks sharing one raster, bench at N GPUs
set -u
N=${1:-2}
TAG=${2:-r2f}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -k 10 600 python -m pytest tests/test_comm_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_comm.log 2>&1; echo "pytest comm rc=$?"; tail -5 gpurun_out/${TAG}_pytest_comm.log
run() {  # name, extra args
  local name=$1; shift
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 "$@" \
    > gpurun_out/${TAG}_bench_${name}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${name}_${N}gpu.err; echo "bench $name rc=$?"; tail -3 gpurun_out/${TAG}_bench_${name}_${N}gpu.err
}
run weak
run strong --scaling strong --no-cpu-baseline
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 \
    > gpurun_out/${TAG}_bench_reference_${N}gpu.json 2> gpurun_out/${TAG}_bench_reference_${N}gpu.err; echo "bench reference rc=$?"
python - <<PY
import json
for name in ("weak", "strong", "reference"):
    f = "gpurun_out/${TAG}_bench_%s_${N}gpu.json" % name
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(name, "n_gpus", d["n_gpus"], "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "scaling", d["scaling"], "e2e", d.get("e2e") and round(d["e2e"]["value"], 2))
        print("   parity", d.get("parity")); print("   cpu", d.get("cpu_baseline")); print("   par", d["config"].get("parallelism"), d["config"].get("comm"))
    except Exception as e:
        print("no bench json", f, e)
PY
