#!/bin/bash
# compute-sanitizer over the NVLink communicator kernels + a small data-parallel step (2+ GPUs, every rank instrumented):
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/sanitize_comm.sh memcheck'
set -u
TOOL="${1:-memcheck}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
export PYTHONUNBUFFERED=1 MINE_B200_MG_QUICK=1 MINE_B200_MG_SMALL=1 MINE_B200_MG_REPLAYS=10
timeout 1100 /usr/local/cuda/bin/compute-sanitizer --tool "$TOOL" --target-processes all --error-exitcode 3 --print-limit 20 \
    --log-file "gpurun_out/sanitizer_comm_${TOOL}_%p.log" \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 tests/multigpu_worker.py \
    > gpurun_out/sanitizer_comm_${TOOL}.out 2>&1
rc=$?
echo "compute-sanitizer --tool $TOOL (comm, $NG ranks) rc=$rc"
grep -h "ERROR SUMMARY" gpurun_out/sanitizer_comm_${TOOL}_*.log | sort | uniq -c
grep -E "RESULT" gpurun_out/sanitizer_comm_${TOOL}.out | cut -c1-400
tail -5 gpurun_out/sanitizer_comm_${TOOL}.out
