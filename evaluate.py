"""Stand-alone evaluation of a checkpoint on the validation split (the reference only evaluates from inside the
training loop, ``synthesis_task.py:476-507``; the metrics and their averaging are the same: L1 / SSIM / PSNR / LPIPS /
log-disparity terms over ``<scene>/images_<ratio>_val``, or over generated pairs for ``synthetic``).

    python evaluate.py --checkpoint_path ws/exp1/checkpoint_latest.pth [--extra_config '{"data.training_set_path": ...}']
                       [--device cpu] [--output metrics.json]

Reads ``params.yaml`` next to the checkpoint (as ``visualizations/image_to_video.py`` does) and prints one JSON line
``{"checkpoint": ..., "num_images": N, "metrics": {name: average}}``.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mine_b200 import config as cfglib  # noqa: E402
from mine_b200.parallel import bootstrap  # noqa: E402
from mine_b200.utils.misc import make_logger  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser(description="Evaluate a checkpoint on the validation split")
    p.add_argument("--checkpoint_path", required=True)
    p.add_argument("--extra_config", default="{}")
    p.add_argument("--device", default=None, choices=[None, "cpu", "cuda"])
    p.add_argument("--output", default=None, help="also write the JSON line to this file")
    args = p.parse_args(argv)

    config = cfglib.load_dumped_config(os.path.join(os.path.dirname(os.path.abspath(args.checkpoint_path)), "params.yaml"),
                                       args.extra_config)
    use_cuda = torch.cuda.is_available() and args.device != "cpu"
    device = torch.device("cuda", 0) if use_cuda else torch.device("cpu")
    logger = make_logger("mine_eval", None)
    config.update({"global_rank": 0, "local_rank": 0, "world_size": 1, "device": device, "tb_writer": None,
                   "logger": logger, "training.pretrained_checkpoint_path": args.checkpoint_path,
                   "engine.resume": False, "engine.cuda_graph": False})

    import train as train_cli
    from synthesis_task import SynthesisTask
    ctx = bootstrap.DistContext(0, 1, 0, device, None)
    _, val_loader = train_cli.get_dataset(config, logger, ctx)
    if len(val_loader) == 0:
        raise FileNotFoundError("the validation split is empty (see input_pipelines/llff/misc/resize_nerf_llff_images.py "
                                "--val_every)")
    task = SynthesisTask(config=config, logger=logger)
    task.run_eval(val_loader)
    n = max(m.count for m in task.val_losses.values())
    metrics = {k: m.avg for k, m in task.val_losses.items()}
    if getattr(task, "lpips", None) is None:
        metrics["lpips_tgt"] = None              # no LPIPS weights on this machine (MINE_LPIPS_WEIGHTS): not measured
    line = json.dumps({"checkpoint": os.path.abspath(args.checkpoint_path), "num_images": int(n), "metrics": metrics})
    print(line)
    if args.output:
        with open(args.output, "w") as f:
            f.write(line + "\n")
    return json.loads(line)


if __name__ == "__main__":
    main()
