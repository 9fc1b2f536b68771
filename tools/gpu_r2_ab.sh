#!/bin/bash
# A/B session: bench.py under a list of environment configurations (one per line in $1), plus the timeline
mkdir -p gpurun_out
T=${2:-r2d}
timeout -k 10 300 python tools/sanitize_target.py visit 3 > gpurun_out/${T}_target_plain.log 2>&1; echo "target plain rc=$?"; tail -2 gpurun_out/${T}_target_plain.log
timeout -k 10 600 python tools/visit_timeline.py c2 > gpurun_out/${T}_timeline_c2.jsonl 2> gpurun_out/${T}_timeline_c2.err; echo "timeline rc=$?"
python - $T <<'PY'
import json,sys
for line in open(f"gpurun_out/{sys.argv[1]}_timeline_c2.jsonl"):
    d=json.loads(line)
    if not d.get("visit",1): continue
    it=d["iterations"]
    print("L%d"%d["level"], "cfg K%d D%d R%d res%d"%(d["cfg"]["supers_per_cta"],d["cfg"]["ring_entries"],d["cfg"]["ring_rounds"],d["cfg"]["resident"]), "total %.1f"%d["total_us"], "pro %.1f min %.1f"%(d["prologue_to_wait_us"],d["min_dt_and_barrier0_arrive_us"]),
          "t0: ring %.1f edge %.1f upd %.1f prod %.1f pro %.1f"%(d["sum_ring_wait_us(thread0)"],d["sum_edge_rounds_us(thread0)"],d["sum_boundary_update_us(thread0)"],d.get("sum_produce_us(thread0)",0),d.get("sum_tile_prologue_us(thread0)",0)),
          "tiles", [x["tiles_us"] for x in it], "bar", [x["barrier_and_issue_us"] for x in it])
PY
while read -r cfg; do
  [ -z "$cfg" ] && continue
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 10 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; rc=$?
  python - "$tag" $T $rc <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[2]}_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "by level", {k:round(v/1e9,2) for k,v in d["flux_edge_updates_per_sec_by_level"].items()}, "roofline", round(d["roofline"]["frac"],3), "sus", round(d.get("sustained",{}).get("ms_per_step",0),4), [ (v["supers_per_cta"],v["ring_entries"],v["ring_rounds"],int(v["resident"])) for v in d["config"]["visit_kernel"]], {k[:9]:round(v["avg_launch_us"],1) for k,v in d["roofline_other"].items()})
except Exception as e: print(sys.argv[1], "rc", sys.argv[3], "parse failed", e)
PY
done < "$1"
