#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 20 --warmup 5 --profile-phases 2>&1 | tail -3 ) > gpurun_out/bench_ours.log
tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log gpurun_out/bench_ours.log
