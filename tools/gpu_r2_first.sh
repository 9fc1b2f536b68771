#!/bin/bash
# round 2, first GPU session: does the visit kernel run, is it right, how fast is it
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout -k 10 300 python tools/sanitize_target.py all 3 > gpurun_out/r2a_target_plain.log 2>&1; echo "target plain rc=$?"; tail -12 gpurun_out/r2a_target_plain.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q --timeout=600 --deselect tests/test_gpu_parity.py::test_c3_class_mesh_matches_the_serial_reference > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest_gpu.log
timeout -k 10 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r2a_bench_c2.json; tail -3 gpurun_out/r2a_bench_c2.err
MGCFD_VISIT=0 timeout -k 10 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c2_stage.json 2> gpurun_out/r2a_bench_c2_stage.err; echo "bench(stage) rc=$?"; cut -c1-1200 gpurun_out/r2a_bench_c2_stage.json
for tool in memcheck synccheck racecheck; do
  timeout -k 10 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py visit 1 > gpurun_out/r2a_sanitizer_${tool}_visit.log 2>&1; echo "$tool rc=$?"; tail -5 gpurun_out/r2a_sanitizer_${tool}_visit.log
done
