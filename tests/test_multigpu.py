"""NVLink all-reduce kernels and the P2P communicator on >= 2 GPUs (one process per GPU, torchrun)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))


def test_p2p_comm_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "multigpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, out.stdout[-2000:] + out.stderr[-3000:]
    res = json.loads(line[-1][len("RESULT "):])
    print(res)
    assert res["small_ok"] and res["mean_ok_p2p"]
    if res.get("multicast"):
        assert res["mean_ok_multimem"]
    assert res["comm_p2p"].startswith("p2p")
    # whole-step check (own comm vs NCCL): meaningful for 2 ranks; with more replicas the receptive-field block's
    # BatchNorm sees >2 nearly identical samples and amplifies bf16 reduction-order noise (both runs are valid)
    if n <= 2:
        assert res["step_grad_cos"] > 0.99
