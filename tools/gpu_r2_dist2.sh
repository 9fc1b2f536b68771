#!/bin/bash
# 2-GPU session: parity of every data plane against one GPU, then bench.py --gpus 2
T=${1:-r2k}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
run() { # name, env..., script args
  name=$1; shift
  timeout -k 10 240 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/dist_check.py > gpurun_out/${T}_dist_${name}.log 2>&1
  echo "dist_check $name rc=$?"; grep -E "dist_check|Error|error" gpurun_out/${T}_dist_${name}.log | tail -8
}
run p2p_stage X=1
if [ "$2" = "all" ]; then
run p2p_visit MGCFD_VISIT=1
run p2p_mixed MGCFD_VISIT=1 MGCFD_VISIT_MAX_NODES=3000
run nccl MGCFD_NO_P2P=1
fi
if [ "$2" = "mixed" ]; then
run p2p_visit MGCFD_VISIT=1
run p2p_mixed MGCFD_VISIT=1 MGCFD_VISIT_MAX_NODES=3000
fi
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-north-star > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/${T}_bench_n2.err
MGCFD_VISIT=1 timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29912 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-north-star > gpurun_out/${T}_bench_n2_visit.json 2> gpurun_out/${T}_bench_n2_visit.err; echo "bench n2 (visit) rc=$?"; tail -3 gpurun_out/${T}_bench_n2_visit.err
timeout -k 10 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
python - $T <<'PY'
import json,sys
for tag in ("n1","n2","n2_visit"):
    try:
        d=json.load(open(f"gpurun_out/{sys.argv[1]}_bench_{tag}.json"))
        print(tag, "ms/step", round(d["ms_per_step"],4), "value", "%.3e"%d["value"], "parity", d.get("parity",{}).get("max_rel_err"), "launches", d["gpu_launches"], "by level", {k:round(v/1e9,2) for k,v in d["flux_edge_updates_per_sec_by_level"].items()})
    except Exception as e: print(tag, "parse failed", e)
PY
