"""Loss curves of the same tiny overfitting run under the three conv executors (semantic check of the
engine's backward: a wrong gradient does not descend like the library reference)."""
import os, sys, copy, torch
sys.path.insert(0, '.')
from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.task import SynthesisTask
base = {"data.img_w": 256, "data.img_h": 128, "mpi.num_bins_coarse": 8, "data.visible_point_count": 64,
        "model.imagenet_pretrained": False, "mpi.fix_disparity": True, "lr.backbone_lr": 2e-4, "lr.decoder_lr": 2e-4}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for mode in ("cudnn_fp32", "cudnn", "tcgen05"):
    os.environ["MINE_B200_CONV"] = mode
    cfg = C.config_for_dataset("llff", dict(base))
    torch.manual_seed(0)
    t = SynthesisTask(cfg, None)
    items = config_batch(cfg, seed=3)
    out = []
    for i in range(steps):
        ld = t.train_step(items)
        if i % 10 == 0 or i == steps - 1:
            out.append("%d:%.4f" % (i, ld["loss"].item()))
    print("%-11s %s" % (mode, "  ".join(out)), flush=True)
