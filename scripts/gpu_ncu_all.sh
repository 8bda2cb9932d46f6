#!/bin/bash
# One ncu --set full capture per kernel family at the headline shapes (single GPU; never a bench value).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_all.sh'   then   python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep
mkdir -p gpurun_out
cap() {   # name, kernel regex, bench_kernels case, launches to skip
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$4" -c 1 -f -o gpurun_out/prof_$1 \
      python scripts/bench_kernels.py --only "$3" --iters 2 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap head_0            "conv_taps"          head_0          4
cap fprop_up_0_1      "conv_taps"          fprop_up_0_1    4
cap fprop_same_2_0    "conv_taps"          fprop_same_2_0  4
cap dgrad_up_0_1      "conv_taps"          dgrad_up_0_1    4
cap wgrad_up_0_1      "wgrad_taps"         wgrad_up_0_1    4
cap wgrad_same_3_0    "wgrad_taps"         wgrad_same_3_0  4
cap bn_act_pad_fwd    "bn_act_pad_fwd"     bn_act_pad_fwd  3
cap bn_act_bwd_reduce "bn_act_bwd_reduce"  bn_act_pad_fwd  3
cap bn_bwd_apply      "bn_bwd_apply"       bn_act_pad_fwd  3
cap render_src_fwd    "render_src_fwd"     render          3
cap render_tgt_fwd    "render_tgt_fwd"     render          3
cap render_tgt_bwd    "render_tgt_bwd"     render          3
cap ssim_fwd          "ssim_fwd"           ssim            3
for f in gpurun_out/prof_*.ncu-rep; do
  n=$(basename "$f" .ncu-rep); python scripts/ncu_summary.py "$f" > "gpurun_out/${n/prof_/ncu_}.txt" 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep | head -20
