#!/bin/bash
T=${2:-r2f}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 -k "visit or cycles_match or baseline_configs or invalid or unstructured" > gpurun_out/${T}_pytest_visit.log 2>&1; echo "pytest(visit) rc=$?"; tail -4 gpurun_out/${T}_pytest_visit.log
bash tools/gpu_r2_ab.sh $1 $T
