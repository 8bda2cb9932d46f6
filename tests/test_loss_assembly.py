"""The fused loss assembly (one stack + addcmul + dot for all scalar terms of all scales, ``task._LossTerms``) must report
the same dictionary and produce the same gradients as the term-by-term arithmetic of the reference
(``synthesis_task.py:329-366``: per-scale sums, cross-scale rule)."""
import os

import pytest
import torch

from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.task import SynthesisTask


@pytest.mark.parametrize("lam1,multi", [(0.5, True), (0.0, False)])
def test_fused_loss_assembly_matches_plain(lam1, multi, monkeypatch):
    shape = {"data.img_w": 128, "data.img_h": 128, "mpi.num_bins_coarse": 4, "data.per_gpu_batch_size": 1,
             "model.imagenet_pretrained": False, "engine.cuda_graph": False, "loss.smoothness_lambda_v1": lam1,
             "training.use_multi_scale": multi}
    cfg = C.config_for_dataset("llff", shape)
    cfg["device"] = torch.device("cpu")
    torch.manual_seed(0)
    task = SynthesisTask(cfg, None)
    items = config_batch(cfg)
    out = {}
    for mode in ("fused", "plain"):
        monkeypatch.setenv("MINE_B200_LOSS_ASSEMBLY", mode)
        torch.manual_seed(1)
        task.set_data(items)
        task.optimizer.zero_grad()
        ld, _ = task.loss_fcn(is_val=False)
        ld["loss"].backward()
        grad = torch.cat([p.grad.reshape(-1) for p in task.decoder.parameters() if p.grad is not None]).clone()
        out[mode] = ({k: float(v.detach()) for k, v in ld.items()}, grad)
    assert set(out["fused"][0]) == set(out["plain"][0])
    for k, ref in out["plain"][0].items():
        assert abs(out["fused"][0][k] - ref) <= 1e-5 * max(1.0, abs(ref)), k
    assert float((out["fused"][1] - out["plain"][1]).norm() / out["plain"][1].norm()) < 1e-5
