"""Convergence parity: the SAME synthetic overfitting run (same initial weights, same data stream, same hyper-
parameters) under three executors - the unmodified reference (``baseline/_ref``, fp32 + cuDNN TF32), this framework
at its default precision (``tf32``) and in the fast mode (``bf16``) - and the loss curves side by side.

    python scripts/train_curves.py --all --steps 2000 --out profiles/loss_curves_r2.json

Every arm runs in its own process (the reference must not share an interpreter with this package).  The reference
arm runs first and writes its initial weights as a checkpoint; the other arms start from that file through the
checkpoint adapter, so all three curves start from identical parameters.  Data: a fixed pool of synthetic source /
target pairs cycled in order; plane disparities are re-sampled every step (stratified sampling, as in training).
"""
import argparse
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INIT = "/tmp/mine_curve_init.pth"
SHAPE = {"w": 384, "h": 256, "planes": 16, "batch": 2, "dataset": "llff"}   # H, W multiples of 128 (upstream decoder)
POOL = 16
KEYS = ("loss", "loss_rgb_tgt", "loss_ssim_tgt", "loss_disp_pt3dsrc", "loss_disp_pt3dtgt", "psnr_tgt")


def batches(device):
    sys.path.insert(0, REPO) if REPO not in sys.path else None
    from mine_b200.data.synthetic import synthetic_batch
    out = []
    for i in range(POOL):
        src, tgt = synthetic_batch(SHAPE["batch"], SHAPE["h"], SHAPE["w"], 256, seed=500 + i)
        out.append(({k: v.to(device) for k, v in src.items()}, {k: v.to(device) for k, v in tgt.items()}))
    return out


def summarise(rows, every):
    """rows: list of dicts of floats per step -> windowed means."""
    curve = []
    for i in range(0, len(rows), every):
        win = rows[i:i + every]
        curve.append({"step": i + len(win), **{k: sum(r[k] for r in win) / len(win) for k in KEYS}})
    return curve


def run_ours(precision, steps, every):
    import torch
    sys.path.insert(0, REPO)
    from mine_b200 import config as C
    from mine_b200.task import SynthesisTask
    extra = {"data.img_w": SHAPE["w"], "data.img_h": SHAPE["h"], "mpi.num_bins_coarse": SHAPE["planes"],
             "data.per_gpu_batch_size": SHAPE["batch"], "model.imagenet_pretrained": False,
             "training.eval_interval": 10 ** 9, "engine.cuda_graph": True, "engine.precision": precision}
    if os.path.exists(INIT):
        extra["training.pretrained_checkpoint_path"] = INIT
    cfg = C.config_for_dataset("llff", extra)
    cfg["device"] = torch.device("cuda:0")
    torch.manual_seed(0)
    task = SynthesisTask(cfg, None)
    data = batches(cfg["device"])
    rows, t0 = [], time.time()
    for i in range(steps):
        ld = task.train_step(data[i % POOL])
        rows.append(torch.stack([ld[k].detach().float().reshape(()) for k in KEYS]))       # no host sync per step
    torch.cuda.synchronize()
    vals = torch.stack(rows).cpu().tolist()
    rows = [dict(zip(KEYS, v)) for v in vals]
    return {"arm": "ours-" + precision, "init": "reference checkpoint" if os.path.exists(INIT) else "own seed 0",
            "seconds": time.time() - t0, "lr": [cfg["lr.backbone_lr"], cfg["lr.decoder_lr"]], "curve": summarise(rows, every)}


def run_reference(steps, every):
    ref_root = os.path.join(REPO, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_root, "synthesis_task.py")):
        return {"arm": "reference", "unavailable": "baseline/_ref is not installed"}
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    import torch
    data = batches(torch.device("cuda:0"))
    from mine_b200.bench.ref_shims import install_import_shims
    install_import_shims()
    sys.path[:] = [ref_root] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO]
    for k in [k for k in sys.modules if k.split(".")[0] in ("utils", "operations", "network", "input_pipelines",
                                                           "synthesis_task", "train")]:
        del sys.modules[k]
    import contextlib
    import io
    import logging
    import warnings
    import numpy as np
    import torch.distributed as dist
    import yaml
    warnings.filterwarnings("ignore")
    if not hasattr(np, "float"):
        np.float = float
    torch.cuda.set_device(0)
    dist.init_process_group(backend="nccl")
    with open(os.path.join(ref_root, "configs", "params_default.yaml")) as f:
        config = yaml.safe_load(f)
    with open(os.path.join(ref_root, "configs", "params_llff.yaml")) as f:
        config.update(yaml.safe_load(f))
    config.update({"data.img_w": SHAPE["w"], "data.img_h": SHAPE["h"], "mpi.num_bins_coarse": SHAPE["planes"],
                   "data.per_gpu_batch_size": SHAPE["batch"], "model.imagenet_pretrained": False,
                   "training.eval_interval": 10 ** 9})
    config["training.gpus"] = [int(s) for s in str(config["training.gpus"]).split(",")]
    config["lr.decay_steps"] = [int(s) for s in str(config["lr.decay_steps"]).split(",")]
    config.update({"current_epoch": 0, "global_rank": 0, "local_rank": 0, "world_size": 1, "tb_writer": None})
    torch.backends.cudnn.benchmark = True
    logger = logging.getLogger("mine_ref_curves")
    logger.addHandler(logging.NullHandler())
    logger.propagate = False
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        from synthesis_task import SynthesisTask
        task = SynthesisTask(config=config, logger=logger)
    # the upstream checkpoint format (DDP "module." prefixes and all): loaded by our checkpoint adapter
    torch.save({"backbone": task.backbone.state_dict(), "decoder": task.decoder.state_dict()}, INIT)
    rows, t0 = [], time.time()
    for i in range(steps):
        task.global_step += 1
        task.set_data(data[i % POOL])
        ld, _ = task.loss_fcn(is_val=False)
        task.optimizer.zero_grad()
        ld["loss"].backward()
        task.optimizer.step()
        rows.append(torch.stack([ld[k].detach().float().reshape(()) for k in KEYS]))
    torch.cuda.synchronize()
    vals = torch.stack(rows).cpu().tolist()
    rows = [dict(zip(KEYS, v)) for v in vals]
    dist.destroy_process_group()
    return {"arm": "reference", "seconds": time.time() - t0, "lr": [config["lr.backbone_lr"], config["lr.decoder_lr"]],
            "curve": summarise(rows, every)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default=None, choices=["reference", "tf32", "bf16"])
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "loss_curves_r2.json"))
    args = ap.parse_args()
    if args.all:
        if os.path.exists(INIT):
            os.remove(INIT)
        results = []
        for arm in ("reference", "tf32", "bf16"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--arm", arm, "--steps", str(args.steps),
                                "--every", str(args.every)], capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("CURVE ")]
            results.append(json.loads(line[-1][6:]) if line else {"arm": arm, "error": (r.stderr or r.stdout)[-1500:]})
            print(arm, "done" if line else "FAILED", flush=True)
        doc = {"what": "same init, same data stream, %d steps, LLFF config at %dx%d N=%d B=%d; windowed means every %d steps"
                       % (args.steps, SHAPE["w"], SHAPE["h"], SHAPE["planes"], SHAPE["batch"], args.every),
               "arms": results}
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(doc, f, indent=1)
        # text table
        arms = [a for a in results if "curve" in a]
        lines = ["step     " + "  ".join("%-22s" % a["arm"] for a in arms)]
        n = min(len(a["curve"]) for a in arms) if arms else 0
        for i in range(n):
            lines.append("%-8d " % arms[0]["curve"][i]["step"] + "  ".join(
                "loss %.4f psnr %5.2f  " % (a["curve"][i]["loss"], a["curve"][i]["psnr_tgt"]) for a in arms))
        txt = "\n".join(lines)
        with open(os.path.splitext(args.out)[0] + ".txt", "w") as f:
            f.write(doc["what"] + "\n" + txt + "\n")
        print(txt)
        return
    res = run_reference(args.steps, args.every) if args.arm == "reference" else run_ours(args.arm, args.steps, args.every)
    print("CURVE " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
