#!/bin/bash
# N-GPU session: parity of the default plane against one GPU, bench.py --gpus N (optionally with the north-star units)
T=${1:-r2w}; N=${2:-4}; NS=${3:---no-north-star}
mkdir -p gpurun_out
timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > gpurun_out/${T}_dist_n${N}.log 2>&1; echo "dist_check n$N rc=$?"; grep -E "^dist_check" gpurun_out/${T}_dist_n${N}.log | tail -7
timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline $NS > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/${T}_bench_n${N}.err | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version"
python - $T $N <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}_bench_n{sys.argv[2]}.json"))
    print("n"+sys.argv[2], "ms/step", round(d["ms_per_step"],4), "value", "%.3e"%d["value"], "parity", d.get("parity",{}).get("max_rel_err"), "launches", d["gpu_launches"])
    for k in ("north_star_c4","north_star_c3"):
        if k in d: print(k, json.dumps(d[k])[:600])
except Exception as e: print("parse failed", e)
PY
