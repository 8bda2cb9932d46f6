#!/bin/bash
# N-GPU validation + measurement: communicator / data-parallel tests, bench (tf32 headline + bf16 line + e2e), step breakdown
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29560"
if [ "${SKIP_SINGLE:-0}" != "1" ]; then
timeout 900 python -m pytest tests -m "gpu and not multigpu" -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2_latest.log 2>&1
echo "single-gpu pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r2_latest.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2_latest.log | head
fi
timeout 1200 python -m pytest tests/test_multigpu.py -q -s -p no:cacheprovider > gpurun_out/pytest_multigpu_r2_n$NG.log 2>&1
echo "multigpu rc=$?"; grep -E "fused_exchange_bit|stat_max|grad_max_rel|exact_cos|identical|ms_152|us_small|passed|failed|Error" gpurun_out/pytest_multigpu_r2_n$NG.log | head -40
timeout 600 $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-render > gpurun_out/bench_n${NG}_full.json 2> gpurun_out/bench_n${NG}_full.err; echo "bench rc=$?"
python - "gpurun_out/bench_n${NG}_full" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(tag + ".json").read().strip().splitlines() if l.startswith("{")][-1])
    print("tf32 %.1f img/s %.2f ms/step e2e %.1f | bf16 %.1f img/s %.2f ms | encoder %s comm %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["fast"].get("value", -1), d["fast"].get("ms_per_step", -1), d["config"].get("encoder"), d["config"].get("comm")))
except Exception as e:
    print(tag, "no result:", e); print(open(tag + ".err").read()[-1500:])
PY
[ "${SKIP_PROFILE:-0}" = "1" ] || timeout 300 $TR scripts/profile_step_multi.py tf32 > gpurun_out/step_breakdown_n${NG}_tf32.txt 2>&1; grep -A40 "^world" gpurun_out/step_breakdown_n${NG}_tf32.txt | head -40
