#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_multigpu.py -q -s --timeout 900 2>&1 | tail -30 ) > gpurun_out/pytest_multigpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 --no-render 2>&1 | tail -3 ) > gpurun_out/bench_ours_n$N.log
grep -E "^\{|passed|failed|Error|error" gpurun_out/pytest_multigpu.log | tail -4 | cut -c1-900; tail -2 gpurun_out/bench_ours_n$N.log
