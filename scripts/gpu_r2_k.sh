#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not multigpu" -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2_latest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r2_latest.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2_latest.log | head
timeout 300 python bench.py --steps 20 --warmup 5 --no-render > gpurun_out/bench_r2_k.json 2> gpurun_out/bench_r2_k.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_k.json').read().strip().splitlines()[-1]);print('tf32',d['value'],d['ms_per_step'],d['e2e']['value'],'bf16',d['fast']['value'],d['fast']['ms_per_step'], 'launches', d['gpu_launches'])"
MINE_B200_GRAD_GATHER=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-render --no-fast --no-e2e > gpurun_out/bench_r2_k_nogather.json 2> gpurun_out/bench_r2_k_nogather.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_k_nogather.json').read().strip().splitlines()[-1]);print('no gather: tf32',d['value'],d['ms_per_step'])"
