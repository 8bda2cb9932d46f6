#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_engine_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2g.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r2g.log
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2g.log | head -20
timeout 200 python scripts/bench_kernels.py --precision tf32 > gpurun_out/kernel_bench_tf32.txt 2>&1; echo "kb tf32 rc=$?"
grep -E "fprop|dgrad|head" gpurun_out/kernel_bench_tf32.txt | grep -E "0_0|0_1|1_0|1_1|head"
timeout 200 python scripts/bench_kernels.py --precision bf16 > gpurun_out/kernel_bench_bf16.txt 2>&1; echo "kb bf16 rc=$?"
grep -E "fprop|dgrad|head" gpurun_out/kernel_bench_bf16.txt | grep -E "0_0|0_1|1_0|1_1|head"
