#!/bin/bash
# one GPU: quick parity subset + two bench runs
T=${1:-r2K}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=600 -k "cycles_match or launches_per_cycle or min_dt_from or transfers or granular or baseline_configs or unstructured" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for v in 1 2; do
timeout -k 10 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err; python - $v gpurun_out/${T}_bench_$v.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{"metric"')][-1])
print("run", sys.argv[1], "ms/step", round(d["ms_per_step"],4), "sustained", round(d["sustained"]["ms_per_step"],4), "launches", d["gpu_launches"], {k: round(v["avg_launch_us"],2) for k,v in d["roofline_other"].items()}, round(d["roofline"]["avg_launch_us"],2))
PY
done
