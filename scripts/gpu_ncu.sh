#!/bin/bash
mkdir -p gpurun_out
for case in head_0 fprop_up_0_1 wgrad_up_0_1 fprop_same_2_0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_taps|wgrad_taps" -s 4 -c 1 -f -o gpurun_out/prof_$case \
     python scripts/bench_kernels.py --only $case --iters 2 > gpurun_out/ncu_$case.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"render_tgt_fwd" -s 3 -c 1 -f -o gpurun_out/prof_render_tgt_fwd \
     python scripts/bench_kernels.py --only render --iters 2 > gpurun_out/ncu_render.log 2>&1
ls -la gpurun_out/*.ncu-rep
