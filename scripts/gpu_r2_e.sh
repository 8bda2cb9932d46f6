#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
if [ "$NG" -ge 2 ]; then
timeout 1200 python -m pytest tests/test_multigpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_multigpu_r2_n$NG.log 2>&1
echo "multigpu rc=$?"; grep -E "cos|exact|rel_err|identical|passed|failed|Error" gpurun_out/pytest_multigpu_r2_n$NG.log | head -40
fi
timeout 900 python scripts/train_curves.py --all --steps 2000 --out gpurun_out/loss_curves_r2.json > gpurun_out/train_curves.log 2>&1; echo "curves rc=$?"
tail -45 gpurun_out/train_curves.log
