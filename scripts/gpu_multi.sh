#!/bin/bash
mkdir -p gpurun_out
( NCCL_DEBUG=WARN timeout 240 python -m pytest tests/test_multigpu.py -q -s --timeout 220 2>&1 | tail -30 ) > gpurun_out/pytest_multigpu8.log
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 20 --warmup 5 --no-render 2>&1 | tail -1 ) > gpurun_out/bench_ours_n8.log
grep -E "^\{|passed|failed" gpurun_out/pytest_multigpu8.log | tail -3 | cut -c1-1500; cut -c1-400 gpurun_out/bench_ours_n8.log
