#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_engine_gpu.py -q -x --timeout 120 2>&1 | tail -60 ) > gpurun_out/conv_tests.log
( timeout 300 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 2>&1 | tail -30 ) > gpurun_out/kernel_tests.log
cat gpurun_out/conv_tests.log; tail -15 gpurun_out/kernel_tests.log
