"""Which data-parallel path reproduces the single-process gradient?  (torchrun, one rank per GPU)"""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
from mine_b200.parallel import bootstrap
ctx = bootstrap.init_distributed()
dev, rank, world = ctx.device, ctx.rank, ctx.world_size
from mine_b200 import config as C
from mine_b200.data.synthetic import synthetic_batch
from mine_b200.task import SynthesisTask
from mine_b200.parallel.comm import Communicator
base = {"data.img_w": 256, "data.img_h": 256, "mpi.num_bins_coarse": 4, "data.visible_point_count": 32,
        "model.imagenet_pretrained": False, "mpi.fix_disparity": True, "data.per_gpu_batch_size": 1,
        "lr.backbone_lr": 0.0, "lr.decoder_lr": 0.0}
items = synthetic_batch(world, 256, 256, 32, seed=0)
mine = tuple({k: v[rank:rank + 1] for k, v in d.items()} for d in items)
res = {}
grads = {}
for kind, mm in (("p2p", "1"), ("p2p", "0"), ("nccl", "0")):
    os.environ["MINE_B200_MULTIMEM"] = mm
    cfg = C.config_for_dataset("llff", dict(base, **{"engine.comm": kind}))
    cfg.update({"device": dev, "global_rank": rank})
    torch.manual_seed(0)
    task = SynthesisTask(cfg, None)
    ld = task.train_step(mine)
    torch.cuda.synchronize(); dist.barrier()
    grads[(kind, mm)] = task.arena.grad.clone()
    res["loss_%s_mm%s" % (kind, mm)] = float(ld["loss"])
    del task
# truth: one process, batch = world, on rank 0 (other ranks idle)
if rank == 0:
    cfg = C.config_for_dataset("llff", dict(base, **{"data.per_gpu_batch_size": world}))
    cfg.update({"device": dev, "global_rank": 0})
    torch.manual_seed(0)
    t = SynthesisTask(cfg, None, comm=Communicator())
    ld = t.train_step(items)
    ref = t.arena.grad
    res["loss_truth"] = float(ld["loss"])
    cos = torch.nn.functional.cosine_similarity
    for k, g in grads.items():
        res["cos_%s_mm%s_vs_truth" % k] = cos(g, ref, dim=0).item()
    res["cos_p2pmm_vs_p2p"] = cos(grads[("p2p", "1")], grads[("p2p", "0")], dim=0).item()
    res["cos_p2p_vs_nccl"] = cos(grads[("p2p", "0")], grads[("nccl", "0")], dim=0).item()
    print("RESULT " + json.dumps(res), flush=True)
dist.barrier(); dist.destroy_process_group()
