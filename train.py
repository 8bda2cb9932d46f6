"""Training entry point (one process per GPU).

Same CLI as upstream ``train.py`` (``--config_path --workspace --version --extra_config
--local_rank``; the dashed ``--local-rank`` of newer launchers and the ``LOCAL_RANK`` env var are
accepted too) and the same outputs under ``<workspace>/<version>/``: ``params.yaml``,
``training.log``, TensorBoard events, ``checkpoint_latest.pth`` / ``checkpoint_%012d.pth``.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py \
        --config_path configs/params_llff.yaml --workspace /tmp/ws --version v0 \
        --extra_config '{"data.training_set_path": "/data/nerf_llff_data"}'

``data.training_set_path: synthetic`` trains on generated source/target pairs of the configured
shape (no dataset needed).
"""
import argparse
import os
import sys

import torch
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mine_b200 import config as cfglib  # noqa: E402
from mine_b200.data.loader import DevicePrefetcher, ShardedSampler  # noqa: E402
from mine_b200.parallel import bootstrap  # noqa: E402
from mine_b200.utils.misc import make_logger, seed_everything  # noqa: E402


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Training")
    p.add_argument("--config_path", default="./params.yaml", type=str)
    p.add_argument("--workspace", type=str, required=True)
    p.add_argument("--version", type=str, required=True)
    p.add_argument("--extra_config", type=str, default="{}")
    p.add_argument("--local_rank", "--local-rank", dest="local_rank", default=None, type=int)
    p.add_argument("--device", default=None, choices=[None, "cpu", "cuda"])
    return p.parse_args(argv)


def get_dataset(config, logger, ctx):
    name = config["data.name"]
    if name not in ("llff", "realestate10k", "flowers", "kitti_raw", "dtu"):
        raise ValueError(f"unknown data.name {name!r}")
    bs = int(config["data.per_gpu_batch_size"])
    root = str(config["data.training_set_path"])
    if root == "synthetic" or root.startswith("synthetic:"):
        from mine_b200.data.synthetic import SyntheticPairs
        n = int(root.split(":")[1]) if ":" in root else 512
        mk = lambda length, seed: SyntheticPairs(length, int(config["data.img_h"]), int(config["data.img_w"]),
                                                 int(config["data.visible_point_count"]), seed=seed)
        train_ds, val_ds = mk(n, 1), mk(max(bs, n // 16), 2)
    elif name == "llff":
        from mine_b200.data.llff import NeRFDataset
        kw = dict(root=root, img_size=(config["data.img_w"], config["data.img_h"]),
                  supervision_count=config["data.num_tgt_views"],
                  visible_points_count=config["data.visible_point_count"],
                  img_pre_downsample_ratio=config["data.img_pre_downsample_ratio"])
        train_ds = NeRFDataset(config, logger, is_validation=False, seed=int(config.get("training.seed", 0)) + ctx.rank, **kw)
        val_ds = NeRFDataset(config, logger, is_validation=True, **kw)
        if len(train_ds) == 0:
            raise FileNotFoundError(
                "no training images under %s/<scene>/%s (COLMAP model in <scene>/sparse/0): create the folder with "
                "input_pipelines/llff/misc/resize_nerf_llff_images.py or set data.img_pre_downsample_ratio to null"
                % (root, train_ds.image_folder))
        if len(val_ds) == 0 and logger:
            logger.info("no validation images under <scene>/%s: evaluation will be skipped" % val_ds.image_folder)
    else:
        raise NotImplementedError(
            f"no loader for {name!r} was ever released upstream; use data.training_set_path=synthetic "
            f"or the LLFF format (eval pair definitions: mine_b200.data.assets)")
    sampler = ShardedSampler(train_ds, ctx.world_size, ctx.rank, shuffle=True, seed=int(config.get("training.seed", 0)))
    workers = max(0, int(config.get("data.num_workers", 0) or 0))     # host-side decode / augmentation workers
    train = DataLoader(train_ds, batch_size=bs, drop_last=True, num_workers=workers, sampler=sampler,
                       collate_fn=train_ds.collate_fn, persistent_workers=workers > 0)
    val = DataLoader(val_ds, batch_size=bs, shuffle=False, drop_last=False, num_workers=0, collate_fn=val_ds.collate_fn)
    return DevicePrefetcher(train, ctx.device), DevicePrefetcher(val, ctx.device)


def main(argv=None):
    args = parse_args(argv)
    config = cfglib.build_config(args.config_path, args.extra_config)
    cfglib.validate_resolution(config)
    ctx = bootstrap.init_distributed(local_rank=args.local_rank, device=args.device)
    config.update({"global_rank": ctx.rank, "local_rank": ctx.local_rank, "world_size": ctx.world_size,
                   "device": ctx.device})
    seed_everything(int(config.get("training.seed", 0)), ctx.rank)

    workspace = os.path.join(args.workspace, args.version)
    logger = None
    if ctx.is_main:
        os.makedirs(workspace, exist_ok=True)
        cfglib.dump_config(config, os.path.join(workspace, "params.yaml"))
        config["log_file"] = "./training.log" if args.workspace.startswith("hdfs") else os.path.join(workspace, "training.log")
        logger = make_logger("mine", config["log_file"])
        logger.info("Training config: {}".format({k: v for k, v in config.items() if k in cfglib.SCHEMA}))
        try:
            from torch.utils.tensorboard import SummaryWriter
            config["tb_writer"] = SummaryWriter(log_dir=workspace)
        except Exception as e:           # tensorboard is optional
            logger.info("TensorBoard disabled: %s" % e)
    config["local_workspace"] = workspace          # every rank knows it (upstream: rank 0 only)
    config["logger"] = logger
    bootstrap.barrier()

    # resume automatically from checkpoint_latest.pth in the workspace
    latest = os.path.join(workspace, "checkpoint_latest.pth")
    if config.get("engine.resume", True) and not config.get("training.pretrained_checkpoint_path") and os.path.exists(latest):
        config["training.pretrained_checkpoint_path"] = latest

    torch.backends.cudnn.benchmark = True
    from synthesis_task import SynthesisTask
    train_loader, val_loader = get_dataset(config, logger, ctx)
    task = SynthesisTask(config=config, logger=logger)
    try:
        task.train(train_loader, val_loader)
    finally:
        if config.get("tb_writer") is not None:
            config["tb_writer"].flush()
        bootstrap.barrier()
        bootstrap.shutdown()


if __name__ == "__main__":
    main()
