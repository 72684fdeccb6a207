#!/bin/bash
set -u
TAG=${1:-leafw}
mkdir -p gpurun_out
timeout -k 10 600 python tools/leaf_check.py --reps 30 --param leaf_impl=2,1 > gpurun_out/${TAG}_leaf_check.txt 2>&1; echo "leaf_check rc=$?"; grep '"k_leaf"' gpurun_out/${TAG}_leaf_check.txt | cut -c1-170
timeout -k 10 900 python -m pytest tests/test_tps_gpu.py tests/test_tiles_gpu.py tests/test_config_scale_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
