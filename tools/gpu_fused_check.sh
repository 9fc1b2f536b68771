#!/bin/bash
# 2-GPU session for the experimental in-kernel halo exchange (MGCFD_P2P_FUSED=1):
#   gpurun --gpus 2 --timeout 300 -- 'bash tools/gpu_fused_check.sh'
# parity against one GPU first (a hang is cut by `timeout`, never left to the box limit), then the C2-sized bench with the
# exchange kernels (baseline) and with the in-kernel exchange.  Outputs in gpurun_out/.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MGCFD_P2P_FUSED=1 timeout 120 $TR --master-port 29531 tools/dist_check.py > gpurun_out/fused_dist2.log 2>&1
echo "fused dist_check rc=$?"; grep dist_check gpurun_out/fused_dist2.log | tail -5
timeout 100 $TR --master-port 29532 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/fused_bench_base.json 2> gpurun_out/fused_bench_base.err
MGCFD_P2P_FUSED=1 timeout 100 $TR --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/fused_bench.json 2> gpurun_out/fused_bench.err
echo "fused bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/fused_bench_base.json", "gpurun_out/fused_bench.json"):
    try:
        d = json.load(open(f)); print(f, "value %.3e  ms/cycle %.3f  launches %d" % (d["value"], d["ms_per_step"], d["gpu_launches"]))
    except Exception as e:
        print(f, "no result:", e)
PY
