#!/usr/bin/env python
"""Headline benchmark: LLFF 384x256, N=32 planes, per-GPU batch 2 - training images/s (whole job).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One rank per GPU (RANK/LOCAL_RANK/WORLD_SIZE from the environment).  Synthetic source/target
pairs of the named shape, random-init weights (no datasets/checkpoints offline); both arms are fed
the same batches.  Timing: W warm-up steps, then exactly K steps inside ONE device-timed region
(CUDA events on the compute stream, barrier + synchronize on both sides), MAX over ranks; a 256 MiB
buffer is rewritten between steps (L2 flush) inside the region.  Rank 0 prints ONE JSON line.

* ``value``   - full optimisation steps (forward, backward, gradient all-reduce, Adam) from
  device-resident inputs.
* ``e2e``     - the same step through the public API a user calls (``SynthesisTask.train_step`` on a
  reference-format batch in *pinned host memory*): every step includes the H2D copies of that
  step's inputs and a D2H read of the loss.
* ``--impl reference`` - the UNMODIFIED upstream code installed in ``baseline/_ref`` driven through
  its own ``SynthesisTask`` methods (``set_data / loss_fcn / backward / optimizer.step``, DDP +
  SyncBN over NCCL), with import shims only for packages missing from the image.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))

BENCH_SHAPES = {
    "llff": {"dataset": "llff", "w": 384, "h": 256, "planes": 32, "batch": 2},
    "realestate": {"dataset": "realestate10k", "w": 384, "h": 256, "planes": 64, "batch": 4},
    "kitti": {"dataset": "kitti_raw", "w": 768, "h": 256, "planes": 32, "batch": 4},
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--shape", default="llff", choices=sorted(BENCH_SHAPES))
    p.add_argument("--precision", default=os.environ.get("MINE_B200_PRECISION", "tf32"), choices=["tf32", "bf16"],
                   help="conv-stack precision of OUR arm: tf32 (default; the reference's class) or bf16 (fast mode)")
    p.add_argument("--no-fast", action="store_true", help="skip the informational bf16 'fast' line")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-render", action="store_true")
    p.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA-graph replay per step")
    p.add_argument("--profile-phases", action="store_true", help="per-phase device times (extra syncs; not a bench value)")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()             # exact PID we started
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:       # region shorter than the sampling period: use the nearest samples
            for ts, line in self.lines[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# shared helpers
# ---------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_batches(shape, n_batches, rank, n_pt=256, pin=True):
    """Reference-format batches in pinned host memory; same generator for both arms."""
    sys.path.insert(0, REPO) if REPO not in sys.path else None
    from mine_b200.data.synthetic import synthetic_batch
    out = []
    for i in range(n_batches):
        out.append(synthetic_batch(shape["batch"], shape["h"], shape["w"], n_pt, seed=1000 * rank + i, pin=pin))
    return out


def tree_bytes(tree):
    import torch
    if torch.is_tensor(tree):
        return tree.numel() * tree.element_size()
    if isinstance(tree, dict):
        return sum(tree_bytes(v) for v in tree.values())
    return sum(tree_bytes(v) for v in tree)


def to_device(items, device):
    src, tgt = items
    mv = lambda d: {k: v.to(device, non_blocking=True) for k, v in d.items()}
    return mv(src), mv(tgt)


def timed_region(step_fn, steps, warmup, device, flush_buf):
    """W untimed steps, then K steps inside one event-timed region; returns elapsed ms (this rank)."""
    import torch
    import torch.distributed as dist
    for i in range(warmup):
        step_fn(i)
    torch.cuda.synchronize(device)
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(steps):
        step_fn(warmup + i)
        flush_buf.add_(1.0)                     # rewrite 256 MiB: evicts L2 between steps
    e1.record()
    torch.cuda.synchronize(device)
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1), t0, time.time()


def max_over_ranks(ms, device):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
DTYPE_TEXT = {
    "tf32": "tf32 (fp32 tensors, TF32 tensor-core convolutions with fp32 accumulation - the reference's precision class; "
            "fp32 BN statistics / render / losses / Adam)",
    "bf16": "bf16 conv stack (bf16 activations and operands, fp32 accumulation, fp32 master weights, fp32 BN statistics "
            "/ render / losses / Adam)",
}


def measure_ours(args, shape, ctx, precision, with_e2e, with_render, with_phases):
    """Build a task at the given conv precision and time it; returns the measurement dict (this rank's view,
    times already max-reduced over ranks)."""
    import torch
    from mine_b200 import config as cfglib
    rank, world, device = ctx.rank, ctx.world_size, ctx.device
    extra = {"data.img_w": shape["w"], "data.img_h": shape["h"], "mpi.num_bins_coarse": shape["planes"],
             "data.per_gpu_batch_size": shape["batch"], "model.imagenet_pretrained": False,
             "training.eval_interval": 10 ** 9, "engine.cuda_graph": not args.no_graph,
             "engine.comm": os.environ.get("MINE_B200_COMM", "p2p"), "engine.precision": precision}
    config = cfglib.config_for_dataset(shape["dataset"], extra)
    config.update({"global_rank": ctx.rank, "local_rank": ctx.local_rank, "world_size": ctx.world_size, "device": device})
    torch.manual_seed(1234 + rank)
    torch.backends.cudnn.benchmark = True
    from mine_b200.task import SynthesisTask
    task = SynthesisTask(config, None)
    from mine_b200.ops import cuda as C

    n_pool = 4
    host_batches = make_batches(shape, n_pool, rank, int(config["data.visible_point_count"]))
    dev_batches = [to_device(b, device) for b in host_batches]
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)
    sampler = ClockSampler(torch.cuda.current_device() if os.environ.get("CUDA_VISIBLE_DEVICES") is None else ctx.local_rank)

    def step_dev(i):
        task.train_step(dev_batches[i % n_pool])

    sampler.start()
    launches0 = C.LAUNCHES["count"]
    ms, t0, t1 = timed_region(step_dev, args.steps, args.warmup, device, flush)
    launches = C.LAUNCHES["count"] - launches0
    # launches during warm-up are included in the counter delta above; subtract them proportionally
    launches = int(round(launches * args.steps / max(args.steps + args.warmup, 1)))
    if task._graph is not None:                  # replays do not tick the Python-side counter
        launches = int(task.launches_per_step) * args.steps
    clocks = sampler.stop(t0, t1)
    ms = max_over_ranks(ms, device)
    out = {"precision": precision, "ms": ms, "launches": launches, "clocks": clocks, "mode": task.runner.mode,
           "comm": task.comm.name, "graph": task._graph is not None,
           "encoder": getattr(getattr(task.runner, "_engine", None), "encoder_mode", "module")}
    if with_e2e:
        h2d = tree_bytes(host_batches[0])

        def step_e2e(i):
            loss = task.train_step(host_batches[i % n_pool])["loss"]
            return float(loss.item())            # D2H read of the step's result

        ms2, _, _ = timed_region(step_e2e, args.steps, 3, device, flush)
        ms2 = max_over_ranks(ms2, device)
        out["e2e"] = {"value": world * shape["batch"] * args.steps / (ms2 / 1e3), "unit": "images/s",
                      "ms_per_step": ms2 / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                      "api": "SynthesisTask.train_step(batch in pinned host memory) + loss.item()"}
    if with_phases:
        task._graph, task._want_graph = None, False
        task.profiler.enabled = True
        for i in range(5):
            step_dev(i)
        out["phases_ms"] = task.profiler.summary()
        task.profiler.enabled = False
    if with_render and rank == 0:
        out["render"] = render_bench_ours(task, config, device)
    if task.grad_sync is not None:
        task.grad_sync.close()
    del task, dev_batches, flush
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    sys.path.insert(0, REPO)
    from mine_b200.parallel import bootstrap
    shape = BENCH_SHAPES[args.shape]
    rank, local_rank, world = dist_env()
    ctx = bootstrap.init_distributed()
    main = measure_ours(args, shape, ctx, args.precision, not args.no_e2e, not args.no_render, args.profile_phases)
    ms = main["ms"]
    result = {
        "metric": "LLFF 384x256 N=32 training images/sec (whole job, device-timed, max over ranks)"
        if args.shape == "llff" else f"{args.shape} training images/sec",
        "impl": "ours",
        "value": world * shape["batch"] * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_TEXT[main["precision"]],
        "data": "synthetic source/target pairs of the named shape, random-init weights",
        "config": {"model": "MINE ResNet-50 encoder + disparity-conditioned MPI decoder (factorised)",
                   "dataset_shape": f"{shape['dataset']} {shape['w']}x{shape['h']} N={shape['planes']}",
                   "global_batch": world * shape["batch"], "per_gpu_batch": shape["batch"], "seq_len": shape["planes"],
                   "parallelism": f"dp{world}", "conv_engine": main["mode"], "encoder": main["encoder"],
                   "precision": main["precision"], "comm": main["comm"], "cuda_graph": main["graph"],
                   "l2": "256 MiB buffer rewritten between steps inside the timed region"},
        "clocks": main["clocks"], "gpu_launches": main["launches"],
    }
    for k in ("e2e", "phases_ms", "render"):
        if k in main:
            result[k] = main[k]
    if not args.no_fast and args.precision != "bf16":
        # extra line: the same step in the reduced-precision fast mode (NOT the headline: lower precision than the reference)
        try:
            fast = measure_ours(args, shape, ctx, "bf16", not args.no_e2e, False, False)
            result["fast"] = {"dtype": DTYPE_TEXT["bf16"], "value": world * shape["batch"] * args.steps / (fast["ms"] / 1e3),
                              "unit": "images/s", "ms_per_step": fast["ms"] / args.steps, "gpu_launches": fast["launches"],
                              "clocks": fast["clocks"], "note": "lower precision than the reference - informational only"}
            if "e2e" in fast:
                result["fast"]["e2e"] = fast["e2e"]
        except Exception as e:          # never lose the headline because of the informational line
            result["fast"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        print(json.dumps(result))
    bootstrap.barrier()
    bootstrap.shutdown()


def render_bench_ours(task, config, device, frames=90, planes=64):
    """Novel-view render latency: B=1, 384x256, N=64 planes, one fused kernel per frame."""
    import torch
    from mine_b200 import geometry as geo
    from mine_b200.ops import api as ops
    h, w = 256, 384
    g = torch.Generator().manual_seed(0)
    mpi = torch.rand(1, planes, h, w, 4, generator=g).to(device)
    disp = torch.linspace(1.0, 0.001, planes, device=device)[None]
    k = geo.fov_intrinsics(h, w, 90.0).to(device)[None]
    kinv = geo.inv3x3(k)
    view = ops.unpack_mpi(mpi)
    poses = torch.eye(4, device=device).repeat(frames, 1, 1)
    poses[:, 0, 3] = torch.linspace(0, -0.16, frames, device=device)
    poses[:, 2, 3] = torch.linspace(0, -0.3, frames, device=device)
    one = torch.ones(1, device=device)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)
    with torch.no_grad():
        for i in range(5):
            task.render_novel_view(view[:, :, :3], view[:, :, 3:], disp, poses[i:i + 1], kinv, k, scale=0, scale_factor=one)
        torch.cuda.synchronize(device)
        tot = 0.0
        for i in range(frames):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            task.render_novel_view(view[:, :, :3], view[:, :, 3:], disp, poses[i:i + 1], kinv, k, scale=0, scale_factor=one)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
    return {"metric": "novel-view render ms/frame (B=1, 384x256, N=64), L2 flushed between frames",
            "ms_per_frame": tot / frames, "frames": frames}


# ---------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------
def run_reference(args):
    ref_root = os.path.join(REPO, "baseline", "_ref")
    rank, local_rank, world = dist_env()
    if not os.path.exists(os.path.join(ref_root, "synthesis_task.py")):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed (run baseline/install_ref.sh)"}))
        return
    # the reference pins every tensor to cuda:0 -> expose exactly one GPU to this process
    os.environ["CUDA_VISIBLE_DEVICES"] = str(local_rank) if os.environ.get("MINE_REF_KEEP_VISIBLE") is None else os.environ["CUDA_VISIBLE_DEVICES"]
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    shape = BENCH_SHAPES[args.shape]
    batches_host = make_batches(shape, 4, rank)      # uses our generator only (data, not model code)
    from mine_b200.bench.ref_shims import install_import_shims
    install_import_shims()
    # from here on nothing of this repository's model/kernels/engine may be imported
    sys.path[:] = [ref_root] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO]
    for k in [k for k in sys.modules if k.split(".")[0] in ("utils", "operations", "network", "input_pipelines",
                                                           "synthesis_task", "train")]:
        del sys.modules[k]
    import logging
    import warnings
    import numpy as np
    import torch
    import torch.distributed as dist
    import yaml
    warnings.filterwarnings("ignore")
    if not hasattr(np, "float"):
        np.float = float
    try:
        torch.cuda.set_device(0)
        dist.init_process_group(backend="nccl")
        device = torch.device("cuda:0")
        with open(os.path.join(ref_root, "configs", "params_default.yaml")) as f:
            config = yaml.safe_load(f)
        with open(os.path.join(ref_root, "configs", {"llff": "params_llff.yaml", "realestate10k": "params_realestate.yaml",
                                                      "kitti_raw": "params_kitti_raw.yaml"}[shape["dataset"]])) as f:
            config.update(yaml.safe_load(f))
        config.update({"data.img_w": shape["w"], "data.img_h": shape["h"], "mpi.num_bins_coarse": shape["planes"],
                       "data.per_gpu_batch_size": shape["batch"], "model.imagenet_pretrained": False,
                       "training.eval_interval": 10 ** 9})
        config["training.gpus"] = [int(s) for s in str(config["training.gpus"]).split(",")]
        config["lr.decay_steps"] = [int(s) for s in str(config["lr.decay_steps"]).split(",")]
        config.update({"current_epoch": 0, "global_rank": dist.get_rank(), "local_rank": local_rank,
                       "world_size": dist.get_world_size(), "tb_writer": None})
        torch.backends.cudnn.benchmark = True
        torch.backends.cudnn.enabled = True
        logger = logging.getLogger("mine_ref_bench")
        logger.addHandler(logging.NullHandler())
        logger.propagate = False
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            from synthesis_task import SynthesisTask
            task = SynthesisTask(config=config, logger=logger)
        mod_file = sys.modules["synthesis_task"].__file__
        assert os.path.abspath(mod_file).startswith(os.path.abspath(ref_root)), mod_file
    except Exception as e:       # the reference cannot run here: say why, exit 0
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {str(e)[:200]}"}))
        return

    dev_batches = [to_device(b, device) for b in batches_host]
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)

    def one_step(items):
        task.global_step += 1
        task.set_data(items)
        loss_dict, _ = task.loss_fcn(is_val=False)
        task.optimizer.zero_grad()
        loss_dict["loss"].backward()
        task.optimizer.step()
        return loss_dict["loss"]

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, t0, t1 = timed_region(lambda i: one_step(dev_batches[i % 4]), args.steps, args.warmup, device, flush)
    clocks = sampler.stop(t0, t1)
    ms = max_over_ranks(ms, device)
    world = dist.get_world_size()
    result = {
        "metric": "LLFF 384x256 N=32 training images/sec (whole job, device-timed, max over ranks)"
        if args.shape == "llff" else f"{args.shape} training images/sec",
        "impl": "reference", "value": world * shape["batch"] * args.steps / (ms / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32 (TF32 cuDNN convs, torch defaults) - the reference has no mixed precision",
        "data": "synthetic source/target pairs of the named shape, random-init weights",
        "config": {"model": "upstream MINE (unmodified, baseline/_ref)",
                   "dataset_shape": f"{shape['dataset']} {shape['w']}x{shape['h']} N={shape['planes']}",
                   "global_batch": world * shape["batch"], "per_gpu_batch": shape["batch"], "seq_len": shape["planes"],
                   "parallelism": f"ddp{world}+syncbn (NCCL)",
                   "l2": "256 MiB buffer rewritten between steps inside the timed region"},
        "clocks": clocks, "gpu_launches": 0,
    }
    if not args.no_e2e:
        ms2, _, _ = timed_region(lambda i: float(one_step(batches_host[i % 4]).item()), args.steps, 2, device, flush)
        ms2 = max_over_ranks(ms2, device)
        result["e2e"] = {"value": world * shape["batch"] * args.steps / (ms2 / 1e3), "unit": "images/s",
                         "ms_per_step": ms2 / args.steps, "h2d_bytes_per_step": tree_bytes(batches_host[0]),
                         "d2h_bytes_per_step": 4, "api": "SynthesisTask.set_data/loss_fcn/backward/step + loss.item()"}
    if not args.no_render and rank == 0:
        try:
            result["render"] = render_bench_reference(task, device)
        except Exception as e:
            result["render"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    if rank == 0:
        print(json.dumps(result))
    dist.barrier()
    dist.destroy_process_group()


def render_bench_reference(task, device, frames=90, planes=64):
    import math
    import torch
    h, w = 256, 384
    g = torch.Generator().manual_seed(0)
    mpi = torch.rand(1, planes, h, w, 4, generator=g).to(device).permute(0, 1, 4, 2, 3).contiguous()
    disp = torch.linspace(1.0, 0.001, planes, device=device)[None]
    f = w * 0.5 / math.tan(math.radians(90.0) * 0.5)
    k = torch.tensor([[f, 0, w * 0.5], [0, f, h * 0.5], [0, 0, 1.0]], device=device)[None]
    kinv = torch.inverse(k)
    poses = torch.eye(4, device=device).repeat(frames, 1, 1)
    poses[:, 0, 3] = torch.linspace(0, -0.16, frames, device=device)
    poses[:, 2, 3] = torch.linspace(0, -0.3, frames, device=device)
    one = torch.ones(1, device=device)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)
    rgb, sig = mpi[:, :, :3], mpi[:, :, 3:]
    with torch.no_grad():
        for i in range(5):
            task.render_novel_view(rgb, sig, disp, poses[i:i + 1], kinv, k, scale=0, scale_factor=one)
        torch.cuda.synchronize(device)
        tot = 0.0
        for i in range(frames):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            task.render_novel_view(rgb, sig, disp, poses[i:i + 1], kinv, k, scale=0, scale_factor=one)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
    return {"metric": "novel-view render ms/frame (B=1, 384x256, N=64), L2 flushed between frames",
            "ms_per_frame": tot / frames, "frames": frames}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
