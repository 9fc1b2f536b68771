#!/bin/bash
# one GPU: transfers / RMS as programmatic dependents, A/B; full capture of the level-0 stage kernel (256-node tiles)
T=${1:-r2H}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=600 -k "cycles_match or launches_per_cycle or min_dt_from or programmatic or granular or baseline_configs" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for v in 0 1 0 1; do
MGCFD_EARLY_RELEASE=$v timeout -k 10 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_er$v.json 2> gpurun_out/${T}_bench_er$v.err; python - $v gpurun_out/${T}_bench_er$v.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{"metric"')][-1])
print("EARLY_RELEASE", sys.argv[1], "ms/step", round(d["ms_per_step"],4), "sustained", round(d["sustained"]["ms_per_step"],4), "launches", d["gpu_launches"])
PY
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_stage_pipe<.int.256" -s 6 -c 3 -o gpurun_out/${T}_prof_stage256 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu_full.log
