"""Where do the small framework launches of one training step come from?  (runs on CPU, no GPU needed)

The step is executed with the conv-engine orchestration routed through the kernel specification (``ops/emu.py``);
every ATen operator dispatched OUTSIDE a region that is one of our kernels on the GPU (engine entry points, the
rendering / loss / optimizer front door) is what becomes "ATen glue" in the CUPTI breakdown
(``profiles/step_breakdown_r1.txt``: ~1.4k launches, 3.9 ms).  Ops are attributed to the innermost repository frame;
operators run by the autograd engine without a Python frame are reported as ``<autograd>``.  View / metadata ops
that launch nothing are filtered out.

    python scripts/count_glue_ops.py [--top 40]
"""
import argparse
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

NO_LAUNCH = {"view", "_unsafe_view", "reshape", "permute", "transpose", "t", "expand", "slice", "select", "unsqueeze",
             "squeeze", "detach", "alias", "as_strided", "unbind", "split", "split_with_sizes", "chunk", "narrow",
             "size", "stride", "sym_size", "sym_stride", "sym_numel", "is_contiguous", "empty", "empty_like",
             "empty_strided", "new_empty", "new_empty_strided", "lift_fresh", "_local_scalar_dense", "unfold", "diagonal",
             "movedim", "view_as", "expand_as", "numel", "dim", "is_same_size", "result_type", "can_cast", "_to_copy_view",
             "new_zeros_meta", "prim_layout", "is_pinned", "storage_offset", "is_nonzero", "_reshape_alias", "unsafe_split",
             "real", "imag", "conj", "resolve_conj", "resolve_neg", "_nested_tensor_size", "ones_like_meta"}


class Counter(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.paused = 0
        self.sites = collections.Counter()
        self.ops = collections.Counter()
        self.kernel_regions = collections.Counter()
        self.backward_ops = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = func.__name__.split(".")[0]
        if self.paused or name in NO_LAUNCH:
            return out
        site = "<autograd>"
        for fr in reversed(traceback.extract_stack()[:-1]):
            fn = fr.filename
            if fn.startswith(REPO) and "/scripts/" not in fn and "count_glue_ops" not in fn:
                site = "%s:%d %s" % (os.path.relpath(fn, REPO), fr.lineno, fr.name)
                break
        self.sites[site] += 1
        self.ops[name] += 1
        if "_train_step_eager" in site or site == "<autograd>":
            self.backward_ops[name] += 1
        return out


def wrap_kernel_region(counter, module, name):
    """Entry points that do not build an autograd graph themselves (the engine's raw kernels, the optimizer)."""
    fn = getattr(module, name)

    def wrapped(*a, **k):
        counter.kernel_regions[name] += 1
        counter.paused += 1
        try:
            return fn(*a, **k)
        finally:
            counter.paused -= 1
    setattr(module, name, wrapped)


class _Region(torch.autograd.Function):
    """Runs ``fn`` as ONE autograd node: its forward and its whole backward are executed with counting paused -
    on the GPU the region is a kernel (or a library module) with a hand-written backward."""

    @staticmethod
    def forward(ctx, counter, fn, box, *flat):
        counter.paused += 1
        try:
            with torch.enable_grad():
                ins = [a.detach().requires_grad_(a.requires_grad) if torch.is_tensor(a) and a.is_floating_point() else a
                       for a in flat]
                out = fn(*ins)
        finally:
            counter.paused -= 1
        if isinstance(out, dict):
            box["keys"], outs = list(out.keys()), list(out.values())
        elif isinstance(out, (tuple, list)):
            box["keys"], outs = None, list(out)
        else:
            box["keys"], outs = "single", [out]
        ctx.counter, ctx.ins, ctx.outs = counter, ins, outs
        return tuple(o.detach() if torch.is_tensor(o) else o for o in outs)

    @staticmethod
    def backward(ctx, *grads):
        counter = ctx.counter
        pairs = [(o, g) for o, g in zip(ctx.outs, grads) if torch.is_tensor(o) and o.requires_grad and g is not None]
        counter.paused += 1
        try:
            if pairs:
                torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        finally:
            counter.paused -= 1
        gin = [a.grad if (torch.is_tensor(a) and a.requires_grad) else None for a in ctx.ins]
        return (None, None, None, *gin)


def wrap_autograd_region(counter, owner, name, label=None):
    fn = getattr(owner, name)
    label = label or name

    def wrapped(*args, **kwargs):
        counter.kernel_regions[label] += 1
        keys = list(kwargs)
        n = len(args)
        call = lambda *flat: fn(*flat[:n], **dict(zip(keys, flat[n:])))
        box = {}
        outs = _Region.apply(counter, call, box, *args, *[kwargs[k] for k in keys])
        if box["keys"] == "single":
            return outs[0]
        if box["keys"] is None:
            return tuple(outs)
        return dict(zip(box["keys"], outs))
    setattr(owner, name, wrapped)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    os.environ["MINE_B200_CONV"] = "tcgen05"
    from mine_b200 import config as C
    from mine_b200.data.synthetic import config_batch
    from mine_b200.ops import conv_engine as E
    from mine_b200.ops import emu
    E.use_emulator(True, torch.float32)
    counter = Counter()
    for name in ("conv_taps", "wgrad_taps", "pack_weights", "bn_act_pad_fwd", "bn_act_bwd_reduce", "bn_bwd_apply",
                 "head_bwd", "bn_res_act_fwd", "bn_res_act_bwd_reduce", "channel_stats", "head_conv_direct"):
        wrap_kernel_region(counter, emu, name)
    # front-door ops that are one forward and one backward kernel on the GPU (here: the PyTorch specification)
    from mine_b200.ops import api
    for name in ("render_src", "render_tgt", "ssim", "masked_l1", "edge_aware_loss", "edge_aware_loss_v2"):
        wrap_autograd_region(counter, api, name)
    import mine_b200.task as T
    from mine_b200.optim import ArenaAdam
    step_fn = ArenaAdam.step

    def adam_step(self):
        counter.kernel_regions["fused_adam"] += 1
        counter.paused += 1
        try:
            return step_fn(self)
        finally:
            counter.paused -= 1
    ArenaAdam.step = adam_step

    shape = {"data.img_w": 128, "data.img_h": 128, "mpi.num_bins_coarse": 4, "data.per_gpu_batch_size": 2,
             "data.visible_point_count": 32, "model.imagenet_pretrained": False}
    cfg = C.config_for_dataset("llff", shape)
    cfg["device"] = torch.device("cpu")
    task = T.SynthesisTask(cfg, None)
    # the encoder trunk is library convolutions + ATen BatchNorm on the GPU (accounted for separately in the CUPTI
    # table): treat it as one region so that its CPU module ops do not drown the glue
    wrap_autograd_region(counter, task.backbone, "forward", "encoder_trunk")
    batch = config_batch(cfg)
    task.train_step(batch)                       # warm-up: caches, lazily built constants
    with counter:
        task.train_step(batch)
    total = sum(counter.sites.values())
    print("kernel regions entered (ours on the GPU): %d  %s" % (sum(counter.kernel_regions.values()),
                                                                 dict(counter.kernel_regions)))
    print("ATen operators outside kernel regions: %d" % total)
    print("  (ops attributed to the line of loss.backward() are autograd-generated backward ops of glue code)")
    print("\nby call site:")
    for site, n in counter.sites.most_common(a.top):
        print("%6d  %s" % (n, site))
    print("\nautograd-generated backward operators of glue code:")
    for op, n in counter.backward_ops.most_common(25):
        print("%6d  %s" % (n, op))
    print("\nby operator:")
    for op, n in counter.ops.most_common(25):
        print("%6d  %s" % (n, op))


if __name__ == "__main__":
    main()
