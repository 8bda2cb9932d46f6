"""Probe for tests/test_kernels_gpu.py::test_cuda_graph_step_matches_eager: gradient cosine of ONE step from identical
weights between (a) two eager runs and (b) eager vs CUDA-graph replay, per encoder mode.  Separates run-to-run noise of
the fp32 atomics (amplified by the random-init network) from a capture bug."""
import copy, os, sys, torch
sys.path.insert(0, '.')
from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.task import SynthesisTask
base = {"data.img_w": 128, "data.img_h": 128, "mpi.num_bins_coarse": 8, "data.visible_point_count": 64,
        "model.imagenet_pretrained": False, "mpi.fix_disparity": True}
for mode in sys.argv[1:] or ["cudnn", "hybrid"]:
    os.environ["MINE_B200_ENCODER"] = mode
    tasks = []
    for graph in (False, False, True):
        cfg = C.config_for_dataset("llff", dict(base, **{"engine.cuda_graph": graph}))
        torch.manual_seed(0)
        t = SynthesisTask(cfg, None)
        if tasks:
            t.arena.data.copy_(tasks[0][0].arena.data)
        tasks.append((t, cfg))
    items = config_batch(tasks[0][1])
    grads, losses = [], []
    for t, _ in tasks:
        l = t.train_step(items)
        grads.append(t.arena.grad.clone()); losses.append(l["loss"].item())
    cos = lambda a, b: torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    print("%s: loss eager %.6f eager2 %.6f graph %.6f | cos(eager, eager2) %.6f  cos(eager, graph) %.6f" %
          (mode, losses[0], losses[1], losses[2], cos(grads[0], grads[1]), cos(grads[0], grads[2])))
