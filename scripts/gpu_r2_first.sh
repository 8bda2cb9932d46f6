#!/bin/bash
# Round 2, first GPU call: validate the TF32 generalisation of every engine kernel and the switched-on variants,
# then measure.  gpurun --timeout 1500 -- 'bash scripts/gpu_r2_first.sh'
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
timeout 900 python -m pytest tests -m "gpu and not multigpu" -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r2.log
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2.log | head -40
timeout 200 python scripts/bench_kernels.py --precision tf32 > gpurun_out/kernel_bench_tf32.txt 2>&1; echo "kb tf32 rc=$?"
timeout 200 python scripts/bench_kernels.py --precision bf16 > gpurun_out/kernel_bench_bf16.txt 2>&1; echo "kb bf16 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2_default.json
MINE_B200_SPARSE=spec MINE_B200_BN_RUNNING=aten MINE_B200_BN_REDUCE=old timeout 200 python bench.py --steps 20 --warmup 5 --no-fast --no-e2e --no-render \
  > gpurun_out/bench_r2_legacy.json 2> gpurun_out/bench_r2_legacy.err; echo "bench legacy rc=$?"
tail -c 600 gpurun_out/bench_r2_legacy.json
timeout 200 python scripts/profile_step.py tf32 > gpurun_out/step_breakdown_tf32.txt 2>&1; echo "profile rc=$?"
head -30 gpurun_out/step_breakdown_tf32.txt
