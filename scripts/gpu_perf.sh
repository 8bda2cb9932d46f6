#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_kernels_gpu.py -q -k "decoder_engine or fused_layer or cuda_graph" --timeout 300 2>&1 | tail -30 ) > gpurun_out/pytest_sel.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-render 2>&1 | tail -2 ) > gpurun_out/bench_ours.log
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 2 --no-graph --no-e2e --no-render > gpurun_out/ncu_bench.log 2>&1 )
tail -12 gpurun_out/pytest_sel.log; cat gpurun_out/bench_ours.log; wc -l gpurun_out/launches.csv
