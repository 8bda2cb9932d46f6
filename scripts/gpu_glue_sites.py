"""Which framework (ATen) operators run in one real GPU training step, by Python call site (TorchDispatchMode; eager).
Counts, not times: almost all of them are ~2 us kernels.  usage: gpu_glue_sites.py [tf32|bf16] [top]"""
import collections, os, sys, traceback, torch
from torch.utils._python_dispatch import TorchDispatchMode
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mine_b200 import config as C
from mine_b200.data.synthetic import config_batch
from mine_b200.task import SynthesisTask
NO_LAUNCH = {"view", "_unsafe_view", "reshape", "permute", "transpose", "t", "expand", "slice", "select", "unsqueeze",
             "squeeze", "detach", "alias", "as_strided", "unbind", "split", "split_with_sizes", "chunk", "narrow",
             "empty", "empty_like", "empty_strided", "new_empty", "new_empty_strided", "lift_fresh", "_local_scalar_dense",
             "unfold", "movedim", "view_as", "expand_as", "_reshape_alias", "unsafe_split", "is_pinned", "record_stream",
             "set_", "resize_", "is_same_size", "_has_compatible_shallow_copy_type"}


class Counter(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.sites = collections.Counter()
        self.ops = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = func.__name__.split(".")[0]
        if name in NO_LAUNCH:
            return out
        site = "<autograd engine>"
        for fr in reversed(traceback.extract_stack()[:-1]):
            fn = fr.filename
            if fn.startswith(REPO) and "/scripts/" not in fn:
                site = "%s:%d %s" % (os.path.relpath(fn, REPO), fr.lineno, fr.name)
                break
        self.sites[(site, name)] += 1
        self.ops[name] += 1
        return out


shape = {"data.img_w": 384, "data.img_h": 256, "mpi.num_bins_coarse": 32, "data.per_gpu_batch_size": 2,
         "model.imagenet_pretrained": False, "engine.cuda_graph": False,
         "engine.precision": sys.argv[1] if len(sys.argv) > 1 else "tf32"}
top = int(sys.argv[2]) if len(sys.argv) > 2 else 70
cfg = C.config_for_dataset("llff", shape)
t = SynthesisTask(cfg, None)
items = tuple({k: v.cuda() for k, v in d.items()} for d in config_batch(cfg))
for _ in range(3):
    t.train_step(items)
torch.cuda.synchronize()
c = Counter()
with c:
    t.train_step(items)
torch.cuda.synchronize()
print("ATen ops with a launch in one step: %d" % sum(c.ops.values()))
print("by operator:", ", ".join("%s %d" % kv for kv in c.ops.most_common(25)))
print("by call site:")
for (site, name), n in c.sites.most_common(top):
    print("%5d  %-22s %s" % (n, name, site))
