"""Kernel-level time table of one eager training step (CUPTI via torch.profiler; not a bench value).
usage: profile_step.py [tf32|bf16]"""
import sys, collections, re, torch
sys.path.insert(0, '.')
from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.task import SynthesisTask
shape = {"data.img_w": 384, "data.img_h": 256, "mpi.num_bins_coarse": 32, "data.per_gpu_batch_size": 2,
         "model.imagenet_pretrained": False, "engine.cuda_graph": False,
         "engine.precision": sys.argv[1] if len(sys.argv) > 1 else "tf32"}
cfg = C.config_for_dataset("llff", shape)
torch.backends.cudnn.benchmark = True
t = SynthesisTask(cfg, None)
items = tuple({k: v.cuda() for k, v in d.items()} for d in config_batch(cfg))
for _ in range(6):
    t.train_step(items)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        t.train_step(items)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace('(anonymous namespace)::', ''); name = re.sub(r'\(.*', '', name); name = re.sub(r'<.*', '', name)[:70]
        agg[name][0] += 1; agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print("precision %s, conv engine %s, encoder %s" % (cfg["engine.precision"], t.runner.mode,
                                                   getattr(t.runner._engine, "encoder_mode", "-")))
print("total device us per step: %.0f" % (tot / 3))
ours = sum(v[1] for k, v in agg.items() if k.startswith("mine::") or " mine::" in k)
lib = sum(v[1] for k, v in agg.items() if ("at::" in k or "cutlass" in k or "cudnn" in k or "xmma" in k or "nvjet" in k))
print("own kernels: %.0f us, ATen/cuDNN/cuBLAS kernels: %.0f us, other (memcpy/memset/...): %.0f us" %
      (ours / 3, lib / 3, (tot - ours - lib) / 3))
for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:60]:
    print("%9.1f us %5.1f  %s" % (us / 3, c / 3, k))
