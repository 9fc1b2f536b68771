#!/bin/bash
T=${1:-r2g}
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 || exit 1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_visit -s 12 -c 4 -o gpurun_out/${T}_prof_visit python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu_full.log
