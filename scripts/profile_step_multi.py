"""Kernel-level time table of one eager N-rank training step on rank 0 (CUPTI via torch.profiler; not a bench value).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/profile_step_multi.py [tf32|bf16]
"""
import sys, collections, re, os, torch
sys.path.insert(0, '.')
from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.parallel import bootstrap
from mine_b200.task import SynthesisTask
ctx = bootstrap.init_distributed()
shape = {"data.img_w": 384, "data.img_h": 256, "mpi.num_bins_coarse": 32, "data.per_gpu_batch_size": 2,
         "model.imagenet_pretrained": False, "engine.cuda_graph": False,
         "engine.precision": sys.argv[1] if len(sys.argv) > 1 else "tf32"}
cfg = C.config_for_dataset("llff", shape)
cfg.update({"global_rank": ctx.rank, "local_rank": ctx.local_rank, "world_size": ctx.world_size, "device": ctx.device})
torch.backends.cudnn.benchmark = True
t = SynthesisTask(cfg, None)
items = tuple({k: v.to(ctx.device) for k, v in d.items()} for d in config_batch(cfg, seed=ctx.rank))
for _ in range(6):
    t.train_step(items)
torch.cuda.synchronize()
bootstrap.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        t.train_step(items)
    torch.cuda.synchronize()
if ctx.rank == 0:
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace('(anonymous namespace)::', ''); name = re.sub(r'\(.*', '', name); name = re.sub(r'<.*', '', name)[:70]
            agg[name][0] += 1; agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print("world %d, precision %s, encoder %s, comm %s" % (ctx.world_size, cfg["engine.precision"],
          getattr(t.runner._engine, "encoder_mode", "-"), t.comm.name))
    print("total device us per step (sum over streams): %.0f" % (tot / 3))
    ours = sum(v[1] for k, v in agg.items() if "mine::" in k)
    comm = sum(v[1] for k, v in agg.items() if "allreduce" in k)
    lib = sum(v[1] for k, v in agg.items() if ("at::" in k or "cutlass" in k or "cudnn" in k or "xmma" in k or "nvjet" in k))
    print("own kernels: %.0f us (of which own collectives %.0f us), ATen/cuDNN/cuBLAS kernels: %.0f us, other: %.0f us" %
          (ours / 3, comm / 3, lib / 3, (tot - ours - lib) / 3))
    for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print("%9.1f us %6.1f  %s" % (us / 3, c / 3, k))
bootstrap.barrier()
bootstrap.shutdown()
