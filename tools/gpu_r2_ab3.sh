#!/bin/bash
T=${2:-r2m}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 -k "visit or cycles_match or baseline_configs or invalid or unstructured or node_kernels or transfers" > gpurun_out/${T}_pytest_visit.log 2>&1; echo "pytest(visit) rc=$?"; tail -4 gpurun_out/${T}_pytest_visit.log
MGCFD_VISIT_WARPS=8 timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 -k "visit or cycles_match or baseline_configs or unstructured" > gpurun_out/${T}_pytest_visit_w8.log 2>&1; echo "pytest(visit, 8 warps) rc=$?"; tail -4 gpurun_out/${T}_pytest_visit_w8.log
MGCFD_VISIT_WARPS=8 timeout -k 10 600 python tools/visit_timeline.py c2 > gpurun_out/${T}_timeline_c2_w8.jsonl 2> gpurun_out/${T}_timeline_c2_w8.err; echo "timeline w8 rc=$?"
python - $T <<'PY'
import json,sys
for line in open(f"gpurun_out/{sys.argv[1]}_timeline_c2_w8.jsonl"):
    d=json.loads(line)
    if not d.get("visit",1): continue
    it=d["iterations"]
    print("w8 L%d"%d["level"], "cfg K%d res%d"%(d["cfg"]["supers_per_cta"],d["cfg"]["resident"]), "total %.1f"%d["total_us"], "pro %.1f min %.1f"%(d["prologue_to_wait_us"],d["min_dt_and_barrier0_arrive_us"]),
          "t0: ring %.1f edge %.1f upd %.1f pro %.1f"%(d["sum_ring_wait_us(thread0)"],d["sum_edge_rounds_us(thread0)"],d["sum_boundary_update_us(thread0)"],d.get("sum_tile_prologue_us(thread0)",0)),
          "tiles", [x["tiles_us"] for x in it], "bar", [x["barrier_and_issue_us"] for x in it])
PY
bash tools/gpu_r2_ab.sh $1 $T
