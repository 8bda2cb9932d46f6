#!/bin/bash
# usage: gpu_ncu_one.sh <case> <kernel regex> [precision]   -> gpurun_out/prof_<case>_<precision>.ncu-rep
mkdir -p gpurun_out
case=$1; regex=$2; prec=${3:-tf32}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 4 -c 1 -f -o gpurun_out/prof_${case}_${prec} \
   python scripts/bench_kernels.py --only $case --iters 2 --precision $prec > gpurun_out/ncu_${case}_${prec}.log 2>&1
ls -la gpurun_out/prof_${case}_${prec}.ncu-rep
