#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_multigpu.py -q -s --timeout 900 2>&1 | tail -30 ) > gpurun_out/pytest_multigpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 --no-render 2>&1 | tail -3 ) > gpurun_out/bench_ours_n$N.log
( MINE_B200_COMM=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 --no-render --no-e2e 2>&1 | tail -2 ) > gpurun_out/bench_ours_nccl_n$N.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29603 bench.py --impl reference --gpus $N --steps 10 --warmup 3 --no-render --no-e2e 2>&1 | tail -2 ) > gpurun_out/bench_ref_n$N.log
grep -E "RESULT|passed|failed|Error|error" gpurun_out/pytest_multigpu.log | tail -8; tail -2 gpurun_out/bench_ours_n$N.log; tail -1 gpurun_out/bench_ours_nccl_n$N.log; tail -1 gpurun_out/bench_ref_n$N.log
