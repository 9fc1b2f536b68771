#!/bin/bash
T=${1:-r2N}
mkdir -p gpurun_out
for v in 12 16 6 12 16; do
MGCFD_EARLY_RELEASE_BLOCKS=$v timeout -k 10 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err; python - $v gpurun_out/${T}_bench_$v.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{"metric"')][-1])
print("blocks/SM", sys.argv[1], "ms/step", round(d["ms_per_step"],4), "sustained", round(d["sustained"]["ms_per_step"],4))
PY
done
