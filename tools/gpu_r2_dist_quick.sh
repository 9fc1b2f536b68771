#!/bin/bash
# N-GPU quick session: guarded parity (small cases + mid-size tets), then the timing probe
T=${1:-r2D}; N=${2:-2}
mkdir -p gpurun_out
for cs in "" tet; do
MGCFD_GUARD=1 timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py $cs > gpurun_out/${T}_dist_n${N}_$cs.log 2>&1; echo "dist_check $cs n$N rc=$?"; grep -E "^dist_check|timed out|rror" gpurun_out/${T}_dist_n${N}_$cs.log | cut -c1-250 | tail -14
done
for wl in c2; do
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/dist_perf.py 200 $wl 2> gpurun_out/${T}_perf_n${N}.err | grep ms_per_cycle | tee -a gpurun_out/${T}_perf.jsonl
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29552 tools/dist_perf.py 200 $wl 2> gpurun_out/${T}_perf_n1.err | grep ms_per_cycle | tee -a gpurun_out/${T}_perf.jsonl
done
