#!/bin/bash
T=${1:-r2t}
N=${2:-2}
mkdir -p gpurun_out
for dbg in ${3:-0 1 2 9 11}; do
  MGCFD_DIST_DEBUG=$dbg timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/dist_perf.py 100 2> gpurun_out/${T}_distperf_$dbg.err | grep ms_per_cycle | tee -a gpurun_out/${T}_distperf.jsonl
done
