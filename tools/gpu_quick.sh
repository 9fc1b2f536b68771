#!/bin/bash
# short GPU session: parity suite, then the A/B probe (tools/exp.py) with the baseline library (if present) and the current one
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
if [ -f mg-cfd-app-plain_b200/libmgcfd_b200_base.so ]; then
  echo "== baseline"; MGCFD_B200_LIB=$PWD/mg-cfd-app-plain_b200/libmgcfd_b200_base.so python tools/exp.py 2>&1 | tee gpurun_out/exp_base.log
fi
echo "== current"; python tools/exp.py 2>&1 | tee gpurun_out/exp.log
