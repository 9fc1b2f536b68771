#!/bin/bash
T=${2:-r2n}
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=600 --deselect tests/test_gpu_parity.py::test_c3_class_mesh_matches_the_serial_reference > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest_gpu.log
while read -r cfg; do
  [ -z "$cfg" ] && continue
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 10 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; rc=$?
  python - "$tag" $T $rc <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[2]}_bench_{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "by level", {k:round(v/1e9,2) for k,v in d["flux_edge_updates_per_sec_by_level"].items()}, "roofline", round(d["roofline"]["frac"],3), "sus", round(d.get("sustained",{}).get("ms_per_step",0),4), "launches", d["gpu_launches"], {k[:9]:round(v["avg_launch_us"],1) for k,v in d["roofline_other"].items()})
except Exception as e: print(sys.argv[1], "rc", sys.argv[3], "parse failed", e)
PY
done < "$1"
