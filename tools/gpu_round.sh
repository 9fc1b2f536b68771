#!/bin/bash
# one GPU session: parity tests, bench, ncu launch list + full capture of the top kernel (B200_PROFILING.md recipe)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench.cu && /tmp/mb > gpurun_out/microbench.log 2>&1; cat gpurun_out/microbench.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_target.py 0 c2 256 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 3 -c 3 -o gpurun_out/prof_flux python tools/ncu_target.py 0 c2 256 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
