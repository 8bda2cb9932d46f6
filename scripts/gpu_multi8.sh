#!/bin/bash
mkdir -p gpurun_out
( NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_multigpu.py -q -s --timeout 600 2>&1 | tail -30 ) > gpurun_out/pytest_multigpu8.log
for N in 8 4; do
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 5 --no-render 2>&1 | tail -1 ) > gpurun_out/bench_ours_n$N.log
done
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 8 --steps 10 --warmup 3 --no-render --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_ref_n8.log
( MINE_B200_COMM=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --steps 20 --warmup 5 --no-render --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_ours_nccl_n8.log
grep -E "^\{|passed|failed" gpurun_out/pytest_multigpu8.log | tail -3 | cut -c1-700; for f in bench_ours_n8 bench_ours_n4 bench_ref_n8 bench_ours_nccl_n8; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['config'].get('comm'), d['config'].get('cuda_graph'), d.get('e2e',{}).get('value'))
except Exception as e: print('$f', 'ERR', open('gpurun_out/$f.log').read()[-400:])
PY
done
