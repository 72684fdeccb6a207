#!/bin/bash
# Round 2, session 3: ncu --set full of k_tri_eig as it is at HEAD (one section point per thread), from a stand-alone 5 000-knot fit
set -u
TAG=${1:-r3x}
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:k_tri_eig" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_tri_eig \
  python tools/fit_check.py default 5000 > gpurun_out/${TAG}_ncu_tri_eig.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/${TAG}_ncu_tri_eig.log | cut -c1-250
