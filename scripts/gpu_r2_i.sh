#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29560"
run_bench() {   # tag, env...
  tag=$1; shift
  env "$@" timeout 300 $TR bench.py --gpus $NG --steps 20 --warmup 5 --no-fast --no-e2e --no-render > gpurun_out/bench_n${NG}_$tag.json 2> gpurun_out/bench_n${NG}_$tag.err
  python - "gpurun_out/bench_n${NG}_$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(tag + ".json").read().strip().splitlines() if l.startswith("{")][-1])
    print("%-32s %.1f img/s  %.2f ms/step  encoder %s comm %s" % (tag, d["value"], d["ms_per_step"], d["config"].get("encoder"), d["config"].get("comm")))
except Exception as e:
    print(tag, "no result:", e); print(open(tag + ".err").read()[-800:])
PY
}
run_bench cudnn X=1
run_bench hybrid MINE_B200_ENCODER=hybrid
timeout 300 $TR scripts/profile_step_multi.py tf32 > gpurun_out/step_breakdown_n${NG}_tf32.txt 2>&1; head -50 gpurun_out/step_breakdown_n${NG}_tf32.txt
