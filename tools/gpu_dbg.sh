#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python tools/dbg_tma.py > gpurun_out/dbg_tma.log 2>&1; echo "rc=$?"; cat gpurun_out/dbg_tma.log | tail -20
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/dbg_tma.py > gpurun_out/dbg_sanitizer.log 2>&1; echo "rc=$?"
grep -v "Host Frame" gpurun_out/dbg_sanitizer.log | head -60
timeout -k 10 900 python tools/ens_check.py both --per-sm 1,2 > gpurun_out/r2e_ens_check.log 2>&1; echo "ens_check rc=$?"; cat gpurun_out/r2e_ens_check.log
