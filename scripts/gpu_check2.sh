#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -150 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 20 --warmup 5 --profile-phases 2>&1 | tail -3 ) > gpurun_out/bench_ours.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-render 2>&1 | tail -1 ) > gpurun_out/bench_ours_nograph.log
grep -E "errs|Error|passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -12; cat gpurun_out/smoke.log gpurun_out/bench_ours.log gpurun_out/bench_ours_nograph.log
