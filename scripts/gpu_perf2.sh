#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m "gpu and not multigpu" -q --timeout 300 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-render --profile-phases 2>&1 | tail -2 ) > gpurun_out/bench_ours.log
grep -E "fused layer grad|passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail; cat gpurun_out/bench_ours.log
