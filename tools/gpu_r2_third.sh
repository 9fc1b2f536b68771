#!/bin/bash
mkdir -p gpurun_out
T=r2c
timeout -k 10 300 python tools/sanitize_target.py visit 3 > gpurun_out/${T}_target_plain.log 2>&1; echo "target plain rc=$?"; tail -6 gpurun_out/${T}_target_plain.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=600 --deselect tests/test_gpu_parity.py::test_c3_class_mesh_matches_the_serial_reference > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest_gpu.log
timeout -k 10 600 python tools/visit_timeline.py c2 > gpurun_out/${T}_timeline_c2.jsonl 2> gpurun_out/${T}_timeline_c2.err; echo "timeline rc=$?"; cat gpurun_out/${T}_timeline_c2.jsonl; tail -3 gpurun_out/${T}_timeline_c2.err
for cfg in "X=1" "MGCFD_PREMIN=0" "MGCFD_VISIT_K=1" "MGCFD_VISIT_D=2 MGCFD_VISIT_R=1" "MGCFD_VISIT_D=4 MGCFD_VISIT_R=1" "MGCFD_VISIT_D=2 MGCFD_VISIT_R=3" "MGCFD_VISIT_D=2 MGCFD_VISIT_R=2" "MGCFD_VISIT_K=2" "MGCFD_VISIT_K=3" "MGCFD_VISIT_K=6"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 10 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $cfg rc=$?"
  python - "$tag" $T <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[2]}_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "by level", {k:round(v/1e9,2) for k,v in d["flux_edge_updates_per_sec_by_level"].items()}, "roofline", round(d["roofline"]["frac"],3), "sus", round(d.get("sustained",{}).get("ms_per_step",0),4), [ (v["supers_per_cta"],v["ring_entries"],v["ring_rounds"],int(v["resident"])) for v in d["config"]["visit_kernel"]], {k[:9]:round(v["avg_launch_us"],1) for k,v in d["roofline_other"].items()})
except Exception as e: print("parse failed", e)
PY
done
tail -3 "gpurun_out/${T}_bench_X_1.err"
