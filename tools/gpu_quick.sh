#!/bin/bash
# short GPU session: parity suite, then the probe (tools/exp.py) with default settings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/exp.py 2>&1 | tee gpurun_out/exp.log
