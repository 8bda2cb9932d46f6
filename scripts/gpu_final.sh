#!/bin/bash
# One GPU-box visit: full GPU test suite, smoke, headline bench, a few kernel timings.  Output -> gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m "gpu and not multigpu" -q --timeout 400 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_ours.log
( timeout 300 python scripts/bench_kernels.py --only bn_act_pad_fwd 2>&1 | grep bn_ ) > gpurun_out/kernel_bn.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_gpu.log | tail -8; cat gpurun_out/smoke.log gpurun_out/kernel_bn.log; python - <<PY
import json; d=json.loads(open('gpurun_out/bench_ours.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['render']['ms_per_frame'], d['gpu_launches'], d['config']['cuda_graph'])
PY
