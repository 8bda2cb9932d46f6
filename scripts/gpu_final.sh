#!/bin/bash
# One GPU-box visit: full GPU test suite, smoke, headline bench, step breakdown.  Output -> gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m "gpu and not multigpu" -q --timeout 400 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_ours.log
( timeout 300 python scripts/profile_step.py 2>&1 | grep -v Warn | tail -48 ) > gpurun_out/step_breakdown.txt
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_gpu.log | tail -8; cat gpurun_out/smoke.log; head -40 gpurun_out/step_breakdown.txt; python - <<PY
import json; d=json.loads(open('gpurun_out/bench_ours.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['render']['ms_per_frame'], d['gpu_launches'], d['config']['cuda_graph'])
PY
