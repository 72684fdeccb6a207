#!/bin/bash
# Round 2, end of round, part 2: the ncu launch list of the bench command (with and without SM partitions)
set -u
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-tiled > gpurun_out/${TAG}_ncu_launch.out 2> gpurun_out/${TAG}_ncu_launch.err; echo "ncu launch list (partitions) rc=$?"
tail -5 gpurun_out/${TAG}_ncu_launch.err; tail -3 gpurun_out/${TAG}_ncu_launch.out | cut -c1-300
wc -l gpurun_out/${TAG}_launches.csv
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches_nopart.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-tiled --param gc_split=-1 > gpurun_out/${TAG}_ncu_launch_nopart.out 2> gpurun_out/${TAG}_ncu_launch_nopart.err; echo "ncu launch list (no partitions) rc=$?"
tail -5 gpurun_out/${TAG}_ncu_launch_nopart.err; wc -l gpurun_out/${TAG}_launches_nopart.csv
