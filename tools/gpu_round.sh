#!/bin/bash
# One GPU session (B200_PROFILING.md recipe): [parity tests,] bench (both arms), then -- each only after its plain run exited 0 --
# the ncu launch list of the bench command and one full capture of the dominant kernel.  Outputs land in gpurun_out/.
# usage: gpu_round.sh [notest]
set -x
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log; fi
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-3000 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_target.py 0 c2 0 3 1 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 2 -c 2 -o gpurun_out/prof_final_c2 python tools/ncu_target.py 0 c2 0 3 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
python tools/assess_compute.py c2 > gpurun_out/assess_compute_c2.jsonl 2> gpurun_out/assess_compute.err; cat gpurun_out/assess_compute_c2.jsonl
