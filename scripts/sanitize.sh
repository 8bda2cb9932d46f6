#!/bin/bash
# compute-sanitizer passes over the kernel tests (the reference has no race / memory checking at all).
# Run on a GPU box, one tool at a time - each multiplies run time by 10-50x, so the kernel-level tests only:
#   gpurun --timeout 900 -- 'bash scripts/sanitize.sh memcheck'
#   tools: memcheck (out-of-bounds / misaligned), racecheck (shared-memory hazards), synccheck (barrier misuse),
#          initcheck (reads of uninitialised global memory)
set -u
TOOL="${1:-memcheck}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 CUDA_LAUNCH_BLOCKING=0
timeout 850 /usr/local/cuda/bin/compute-sanitizer --tool "$TOOL" --error-exitcode 3 --print-limit 20 \
    --log-file "gpurun_out/sanitizer_${TOOL}.log" \
    python -m pytest tests/test_kernels_gpu.py tests/test_conv_engine_gpu.py -x -q \
        -k "render_src or render_tgt or ssim or masked_l1 or smoothness or fused_adam or conv_same_fprop or conv_up_fprop or dgrad_and_wgrad_same or dgrad_and_wgrad_up or bn_act_pad or fused_layer"
rc=$?
echo "compute-sanitizer --tool $TOOL rc=$rc (exit code 3 = errors reported)"
tail -15 "gpurun_out/sanitizer_${TOOL}.log"
exit $rc
