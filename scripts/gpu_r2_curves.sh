#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 1500 python scripts/train_curves.py --all --steps 2000 --out gpurun_out/loss_curves_r2.json > gpurun_out/train_curves.log 2>&1; echo "curves rc=$?"
tail -45 gpurun_out/train_curves.log
