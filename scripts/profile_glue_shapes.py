"""ATen operators of one eager training step grouped by input shapes, by device time (CUPTI; not a bench value).
usage: profile_glue_shapes.py [tf32|bf16] [op substring ...]"""
import sys, collections, torch
sys.path.insert(0, '.')
from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.task import SynthesisTask
shape = {"data.img_w": 384, "data.img_h": 256, "mpi.num_bins_coarse": 32, "data.per_gpu_batch_size": 2,
         "model.imagenet_pretrained": False, "engine.cuda_graph": False,
         "engine.precision": sys.argv[1] if len(sys.argv) > 1 else "tf32"}
want = sys.argv[2:] or ["copy_", "add_", "fill_", "add", "mul"]
cfg = C.config_for_dataset("llff", shape)
torch.backends.cudnn.benchmark = True
t = SynthesisTask(cfg, None)
items = tuple({k: v.cuda() for k, v in d.items()} for d in config_batch(cfg))
for _ in range(5):
    t.train_step(items)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    t.train_step(items)
    torch.cuda.synchronize()
for w in want:
    rows = []
    for ka in prof.key_averages(group_by_input_shape=True):
        if ka.key != "aten::" + w:
            continue
        dt = getattr(ka, "self_device_time_total", 0) or 0
        if dt > 0:
            rows.append((dt, ka.count, str(ka.input_shapes)[:110]))
    print("== aten::%s: %.0f us" % (w, sum(r[0] for r in rows)))
    for dt, c, shp in sorted(rows, reverse=True)[:14]:
        print("  %8.1f us %3d  %s" % (dt, c, shp))
