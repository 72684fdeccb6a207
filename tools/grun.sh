#!/bin/bash
# gpurun wrapper: rebuild the library first (the GPU box runs the .so that travels with the snapshot, not the sources)
set -e
cd "$(dirname "$0")/.."
python -m machisplin_b200.build > /dev/null
make -s -C oracle/c
exec /usr/local/graft/bin/gpurun "$@"
