#!/bin/bash
# one GPU: the tests the last session did not reach, and the distributed code path on ONE rank (no peers): its cost without any communication
T=${1:-r2B}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=900 -k "c3_class or outside_its_arrays or guard_zones" > gpurun_out/${T}_pytest_rest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest_rest.log
for dbg in 0 9; do
MGCFD_DIST_DEBUG=$dbg timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29551 tools/dist_perf.py 200 c2 2> gpurun_out/${T}_onerank_$dbg.err | grep ms_per_cycle | tee -a gpurun_out/${T}_onerank.jsonl
done
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_onerank_launches.csv python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29552 tools/dist_perf.py 3 c2 > gpurun_out/${T}_onerank_ncu.log 2>&1; echo "ncu rc=$?"
