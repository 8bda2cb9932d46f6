#!/bin/bash
# full single-GPU test tier + 2-GPU communicator / data-parallel equivalence + convergence curves
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not multigpu" -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2c.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2c.log
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2c.log | head -40
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_multigpu_r2_n$NG.log 2>&1
  echo "multigpu rc=$?"; tail -60 gpurun_out/pytest_multigpu_r2_n$NG.log
fi
