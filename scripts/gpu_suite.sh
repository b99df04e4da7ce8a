#!/bin/bash
# Runs every GPU test file in its own process (a sticky CUDA error then only costs that file),
# each under a timeout, logging to gpurun_out/.  Usage: scripts/gpu_suite.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_geometry_gpu.py tests/test_network_gpu.py tests/test_bifusion.py tests/test_plus.py tests/test_tools_gpu.py tests/test_model_gpu.py tests/test_vitl_gpu.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -x --timeout 600 "$@" > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f -> exit $r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
