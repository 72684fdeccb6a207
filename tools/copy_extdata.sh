#!/bin/bash
# Build container only: copy the reference's three bundled rasters into the git-ignored baseline/_ref/extdata so that they travel
# to the GPU box with the gpurun snapshot (tools/ens_check.py real).  Nothing in tests/, bench.py or smoke() reads them.
set -e
mkdir -p baseline/_ref/extdata
cp /root/reference/inst/extdata/{alt,slope,TWI}.tif baseline/_ref/extdata/
