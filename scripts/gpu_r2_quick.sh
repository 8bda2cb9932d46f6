#!/bin/bash
# quick loop: conv-engine tests, tf32 kernel bench, short bench (no bf16 line)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_optin_kernels_gpu.py -q -p no:cacheprovider -x > gpurun_out/pytest_gpu_quick.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_quick.log
timeout 200 python scripts/bench_kernels.py --precision tf32 > gpurun_out/kernel_bench_tf32.txt 2>&1; echo "kb tf32 rc=$?"
grep -E "ms " gpurun_out/kernel_bench_tf32.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-render --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]);print('tf32',d['value'],d['ms_per_step'],'bf16',d['fast']['value'],d['fast']['ms_per_step'])"
