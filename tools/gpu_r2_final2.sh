#!/bin/bash
# closing one-GPU session: whole GPU suite, both bench arms, launch list
T=${1:-r2P}
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${T}_bench_c2.json; tail -2 gpurun_out/${T}_bench_c2.err
timeout -k 10 400 python bench.py --impl reference --steps 3 --warmup 1 --with-serial > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${T}_bench_ref.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout -k 10 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-250 gpurun_out/${T}_bench_c3.json
