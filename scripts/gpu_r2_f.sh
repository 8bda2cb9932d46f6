#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_encoder_engine_gpu.py tests/test_kernels_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2f.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2f.log
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2f.log | head -40
timeout 200 python scripts/bench_kernels.py --precision tf32 > gpurun_out/kernel_bench_tf32.txt 2>&1; echo "kb tf32 rc=$?"
grep -E "0_0|0_1|1_0|1_1|head" gpurun_out/kernel_bench_tf32.txt
timeout 200 python scripts/bench_kernels.py --precision bf16 > gpurun_out/kernel_bench_bf16.txt 2>&1; echo "kb bf16 rc=$?"
grep -E "0_0|0_1|1_0|1_1|head" gpurun_out/kernel_bench_bf16.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-render > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_f.json').read().strip().splitlines()[-1]);print('tf32',d['value'],d['ms_per_step'],d['e2e']['value'],'bf16',d['fast']['value'],d['fast']['ms_per_step'])"
