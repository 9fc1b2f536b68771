#!/bin/bash
# one-GPU session: the whole GPU test suite, both bench arms, the ncu launch list and one full capture of the dominant kernel
T=${1:-r2A}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout -k 10 1700 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest_gpu.log
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/${T}_bench_c2.json; tail -2 gpurun_out/${T}_bench_c2.err
timeout -k 10 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_ref.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_pipe -s 30 -c 3 -o gpurun_out/${T}_prof_stage python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu_full.log
