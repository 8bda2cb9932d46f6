#!/bin/bash
# First GPU call of the next round: validate every opt-in path written without hardware access and measure each
# switch separately and together.  gpurun --timeout 600 -- 'bash scripts/gpu_round2_kickoff.sh'
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
MINE_B200_TEST_OPTIN=1 timeout 200 python -m pytest tests/test_optin_kernels_gpu.py tests/test_encoder_engine_gpu.py -q \
    > gpurun_out/t_optin.log 2>&1; echo "opt-in tests rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_optin.log | head -30
timeout 60 python scripts/bench_optin.py 2>&1 | tee gpurun_out/bench_optin_default.txt
MINE_B200_HEAD=direct MINE_B200_BN_REDUCE=v2 timeout 60 python scripts/bench_optin.py 2>&1 | tee gpurun_out/bench_optin_variants.txt
run_bench() {   # tag, env...
  tag=$1; shift
  env "$@" timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % tag).read().strip().splitlines()[-1])
    print("%-12s %.1f img/s  %.2f ms/step" % (tag, d["value"], d["ms_per_step"]))
except Exception as e:
    print(tag, "no result:", e); print(open("gpurun_out/bench_%s.err" % tag).read()[-800:])
PY
}
run_bench default X=1
run_bench head MINE_B200_HEAD=direct
run_bench bnv2 MINE_B200_BN_REDUCE=v2
run_bench sparse MINE_B200_SPARSE=fused
run_bench running MINE_B200_BN_RUNNING=fused
run_bench all MINE_B200_HEAD=direct MINE_B200_BN_REDUCE=v2 MINE_B200_SPARSE=fused MINE_B200_BN_RUNNING=fused
run_bench enc_engine MINE_B200_ENCODER=tcgen05
run_bench enc_splitk MINE_B200_ENCODER=tcgen05 MINE_B200_SPLITK=1
