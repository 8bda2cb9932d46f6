#!/bin/bash
run() { echo "== $1"; shift; env "$@" timeout 120 python scripts/bench_kernels.py --only head_0 --iters 5 2>&1 | grep head_0; env "$@" timeout 120 python scripts/bench_kernels.py --only fprop_up_0_1 --iters 5 2>&1 | grep fprop; }
run base X=1
run no_epilogue MINE_CONV_DBG=8
run no_mma MINE_CONV_DBG=4
run no_wtma MINE_CONV_DBG=1
run no_tma MINE_CONV_DBG=3
run no_tma_no_mma MINE_CONV_DBG=7
run nothing MINE_CONV_DBG=15
run ipb1 MINE_CONV_IPB=1 MINE_CONV_STAGES=8
run ipb9_st2 MINE_CONV_IPB=9 MINE_CONV_STAGES=2
