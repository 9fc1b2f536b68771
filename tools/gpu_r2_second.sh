#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=600 --deselect tests/test_gpu_parity.py::test_c3_class_mesh_matches_the_serial_reference > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
timeout -k 10 600 python tools/visit_timeline.py c2 > gpurun_out/r2b_timeline_c2.jsonl 2> gpurun_out/r2b_timeline_c2.err; echo "timeline rc=$?"; cat gpurun_out/r2b_timeline_c2.jsonl; tail -3 gpurun_out/r2b_timeline_c2.err
MGCFD_VISIT_R=1 timeout -k 10 600 python tools/visit_timeline.py c2 0 3 > gpurun_out/r2b_timeline_c2_R1.jsonl 2>&1; cat gpurun_out/r2b_timeline_c2_R1.jsonl
for cfg in "X=1" "MGCFD_VISIT_R=1" "MGCFD_VISIT_R=2" "MGCFD_VISIT_K=2" "MGCFD_VISIT_RESIDENT=0"; do
  env $cfg timeout -k 10 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_$cfg.json 2> gpurun_out/r2b_bench_$cfg.err; echo "bench $cfg rc=$?"
  python - "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/r2b_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "by level", {k:round(v/1e9,2) for k,v in d["flux_edge_updates_per_sec_by_level"].items()}, "roofline", round(d["roofline"]["frac"],3), "sus", d.get("sustained",{}).get("ms_per_step"), [ (v["supers_per_cta"],v["ring_rounds"],v["resident"]) for v in d["config"]["visit_kernel"]])
except Exception as e: print("parse failed", e)
PY
done
tail -3 "gpurun_out/r2b_bench_X=1.err"
timeout -k 10 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2b_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2b_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_visit -s 12 -c 6 -o gpurun_out/r2b_prof_visit python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2b_ncu_full.log
