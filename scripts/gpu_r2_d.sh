#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
timeout 1200 python -m pytest tests/test_multigpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_multigpu_r2_n$NG.log 2>&1
echo "multigpu rc=$?"; grep -E "cos|exact|audit|rel_err|identical|soak|passed|failed|Error" gpurun_out/pytest_multigpu_r2_n$NG.log | head -60
