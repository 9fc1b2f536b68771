#!/bin/bash
# closing 2-GPU session: the 2-GPU tests of the suite, guarded parity, bench --gpus 2
T=${1:-r2Q}
mkdir -p gpurun_out
timeout -k 10 500 python -m pytest tests/test_gpu_dist.py tests/test_driver.py -m gpu -x -q --timeout=400 > gpurun_out/${T}_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_dist.log
MGCFD_GUARD=1 timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > gpurun_out/${T}_dist_n2.log 2>&1; echo "dist_check rc=$?"; grep -E "^dist_check (PASS|FAIL)|damaged on all ranks: [1-9]" gpurun_out/${T}_dist_n2.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-north-star > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench rc=$?"
python - gpurun_out/${T}_bench_n2.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{"metric"')][-1])
print("n2 ms/step", round(d["ms_per_step"],4), "value %.3e" % d["value"], "parity", d["parity"]["max_rel_err"])
PY
