#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_encoder_engine_gpu.py tests/test_optin_kernels_gpu.py tests/test_kernels_gpu.py tests/test_task_vs_reference_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2b.log
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2b.log | head -40
timeout 200 python scripts/bench_kernels.py --precision tf32 > gpurun_out/kernel_bench_tf32.txt 2>&1; echo "kb tf32 rc=$?"
grep wgrad gpurun_out/kernel_bench_tf32.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-fast --no-render > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_b.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
