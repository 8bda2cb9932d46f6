"""Which framework (ATen) operators still run inside a training step, by call site and device time (CUPTI via
torch.profiler with Python stacks; not a bench value).  usage: profile_glue.py [tf32|bf16]"""
import sys, collections, torch
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
for _ in range(5):
    t.train_step(items)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    t.train_step(items)
    torch.cuda.synchronize()
rows = []
for ka in prof.key_averages(group_by_stack_n=12):
    if not ka.key.startswith("aten::"):
        continue
    dt = getattr(ka, "self_device_time_total", 0) or 0
    if dt <= 0:
        continue
    site = "?"
    for fr in (ka.stack or []):
        if ("/mine_b200/" in fr or "/repo/task" in fr) and "profile_glue" not in fr:
            site = fr.split("/mine_b200/")[-1][:70]
            break
    rows.append((dt, ka.count, ka.key, site))
agg = collections.defaultdict(lambda: [0, 0.0])
for dt, c, k, site in rows:
    agg[(k, site)][0] += c
    agg[(k, site)][1] += dt
tot = sum(v[1] for v in agg.values())
print("ATen self device time per step: %.0f us in %d op calls" % (tot, sum(v[0] for v in agg.values())))
for (name, site), (c, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:80]:
    print("%8.1f us %4d  %-34s %s" % (us, c, name, site))
