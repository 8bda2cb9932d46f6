#!/bin/bash
# One short GPU call: re-validate the shared conv kernels after the channel-block / strided-wgrad extension,
# run the opt-in encoder-engine tests, and time the step with both encoder modes.
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_conv_engine_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv-engine tests rc=$?"
tail -3 gpurun_out/t_conv.log
MINE_B200_TEST_ENCODER=1 timeout 150 python -m pytest tests/test_encoder_engine_gpu.py -q > gpurun_out/t_enc.log 2>&1; echo "encoder tests rc=$?"
grep -E "passed|failed|Error|assert " gpurun_out/t_enc.log | tail -25
timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
python - <<'PY'
import json
for n in ("default",):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1]); print(n, d["value"], d["ms_per_step"], d["gpu_launches"])
    except Exception as e: print(n, "no result", e)
PY
MINE_B200_ENCODER=tcgen05 timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_enc.json 2> gpurun_out/bench_enc.err; echo "bench encoder-engine rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_enc.json").read().strip().splitlines()[-1]); print("encoder-engine", d["value"], d["ms_per_step"], d["gpu_launches"])
except Exception as e:
    print("encoder-engine no result", e); print(open("gpurun_out/bench_enc.err").read()[-1500:])
PY
MINE_B200_ENCODER=hybrid timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err; echo "bench hybrid rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_hybrid.json").read().strip().splitlines()[-1]); print("hybrid", d["value"], d["ms_per_step"], d["gpu_launches"])
except Exception as e:
    print("hybrid no result", e); print(open("gpurun_out/bench_hybrid.err").read()[-1500:])
PY
