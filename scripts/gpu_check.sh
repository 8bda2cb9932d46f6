#!/bin/bash
# One GPU-box visit: tests, smoke, both bench arms, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -20 ) > gpurun_out/smoke.log
( timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -5 ) > gpurun_out/bench_ref.log
( timeout 900 python bench.py --steps 20 --warmup 5 --profile-phases 2>&1 | tail -5 ) > gpurun_out/bench_ours.log
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-render > gpurun_out/ncu_bench.log 2>&1 )
tail -c 1500 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log gpurun_out/bench_ref.log gpurun_out/bench_ours.log
