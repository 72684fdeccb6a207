#!/bin/bash
# round 2, session 3: why do polynomial exponentials slow k_ens_svm_tma down?  ncu --set full with all exponentials on MUFU.EX2 (svm_poly 1) and 2/8 on the FMA pipe (3)
set -u
TAG=${1:-r3b}
mkdir -p gpurun_out
for P in 1 3; do
  timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:k_ens_svm_tma" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_svm_poly$P \
    python tools/ens_check.py synthetic --kept v --svm 1 --poly $P --reps 1 > gpurun_out/${TAG}_ncu_svm_poly$P.log 2>&1; echo "ncu poly $P rc=$?"
  tail -2 gpurun_out/${TAG}_ncu_svm_poly$P.log
done
ls -la gpurun_out/${TAG}_prof_*
