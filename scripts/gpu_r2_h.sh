#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
run_bench() {   # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-fast --no-e2e --no-render > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % tag).read().strip().splitlines()[-1])
    print("%-14s %.1f img/s  %.2f ms/step  launches %s encoder %s" % (tag, d["value"], d["ms_per_step"], d["gpu_launches"], d["config"].get("encoder")))
except Exception as e:
    print(tag, "no result:", e); print(open("gpurun_out/bench_%s.err" % tag).read()[-800:])
PY
}
run_bench enc_cudnn X=1
run_bench enc_engine MINE_B200_ENCODER=tcgen05
run_bench enc_splitk MINE_B200_ENCODER=tcgen05 MINE_B200_SPLITK=1
run_bench enc_hybrid MINE_B200_ENCODER=hybrid
MINE_B200_ENCODER=tcgen05 MINE_B200_SPLITK=1 timeout 200 python -m pytest tests/test_encoder_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
