#!/bin/bash
set -u
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_geotiff_gpu.py tests/test_tps_gpu.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout -k 10 900 python tools/tiff_check.py > gpurun_out/${TAG}_tiff_check.txt 2>&1; echo "tiff_check rc=$?"; cat gpurun_out/${TAG}_tiff_check.txt | tail -6
