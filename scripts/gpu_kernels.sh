#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m "gpu and not multigpu" -q --timeout 300 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log
( timeout 600 python scripts/bench_kernels.py 2>&1 | grep -v Warn | tail -50 ) > gpurun_out/kernel_bench.log
( timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_ours.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -5; cat gpurun_out/kernel_bench.log gpurun_out/bench_ours.log
