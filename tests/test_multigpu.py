"""NVLink all-reduce kernels and the P2P communicator on >= 2 GPUs (one process per GPU, torchrun): kernel-level
equality with NCCL, an in-step audit of every collective of a training step at the bench shape, data-parallel
equivalence (own communicator == NCCL communicator == one process on the N x batch) and a graph-replay soak."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def run_worker(n, env=None, timeout=1500):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "multigpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, out.stdout[-2000:] + out.stderr[-3000:]
    return json.loads(line[-1][len("RESULT "):])


def check(res):
    n = res["world"]
    assert res["small_ok"] and res["mean_ok_p2p"]
    assert res["fused_exchange_bit_identical"], res
    if res.get("multicast"):
        assert res["mean_ok_multimem"]
    assert res["comm_p2p"].startswith("p2p")
    # every collective of a real step equals NCCL's result for the same input (fp32 summation-order noise only)
    for tag in ("p2p", "multimem"):
        a = res.get("audit_" + tag)
        if a is None:
            continue
        assert a["stat_collectives"] >= 100 and a["grad_buckets"] >= 4, a
        assert a["stat_max_rel_err"] < 1e-5, (tag, a)
        assert a["grad_max_rel_err"] < 1e-5, (tag, a)
        assert a["grads_identical_across_ranks"], (tag, a)
    # Whole-step gradient cosines WITH the kernels (own vs NCCL communicator, vs a repeated NCCL run, vs one process)
    # are reported but not asserted: TF32 rounding noise through this random-init network is O(1) on the gradient
    # direction (two NCCL runs are as far apart as anything else; scripts/conditioning_probe.py), while the audit above
    # proves the collectives themselves are exact.
    # data-parallel equivalence in exact arithmetic (fp32 specification kernels): N ranks == one process on N x batch
    assert res["exact_cos_p2p_vs_single_process"] > 0.9995, res
    if "exact_cos_multimem_vs_single_process" in res:
        assert res["exact_cos_multimem_vs_single_process"] > 0.9995, res
    s = res["soak"]
    assert s["graph"] and s["params_bit_identical_across_ranks"] and s["params_finite"], s
    assert n >= 2


def test_p2p_comm_and_data_parallel_step_match_nccl():
    n = min(torch.cuda.device_count(), 8)
    res = run_worker(n)
    print(json.dumps(res, indent=1))
    check(res)
