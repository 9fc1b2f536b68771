#!/bin/bash
# final one-GPU session of the round: the whole GPU suite, both bench arms, assess-compute sweep (+ counters), launch list, full capture
T=${1:-r2G}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest_gpu.log
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench_c2.json; tail -2 gpurun_out/${T}_bench_c2.err
timeout -k 10 400 python bench.py --impl reference --steps 3 --warmup 1 --with-serial > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${T}_bench_ref.json
for m in c2 tet; do timeout -k 10 300 python tools/assess_compute.py $m > gpurun_out/${T}_assess_$m.jsonl 2> gpurun_out/${T}_assess_$m.err; echo "assess $m rc=$?"; done
cat gpurun_out/${T}_assess_c2.jsonl | cut -c1-260
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_flux_assess|k_indirect_rw|k_flux_atomic" -c 60 --csv --log-file gpurun_out/${T}_assess_ncu.csv python tools/assess_compute.py c2 > gpurun_out/${T}_assess_ncu.log 2>&1; echo "ncu assess rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_pipe -s 30 -c 3 -o gpurun_out/${T}_prof_stage python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu_full.log
