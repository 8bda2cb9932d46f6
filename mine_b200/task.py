"""SynthesisTask: trainer / evaluator / renderer facade (public API parity).

Reference: ``synthesis_task.py:63-670``.  Same constructor, attributes and method names
(``set_data, network_forward, mpi_predictor, loss_fcn, loss_fcn_per_scale, render_novel_view, train,
train_epoch, run_eval, log_training, log_val, compute_scale_factor, init_data``) and the same batch
and checkpoint formats - on a different execution stack:

* device-agnostic (no ``cuda:0`` pinning), zero host synchronisations inside a step (closed-form
  inverses, no ``.item()`` unless logging), so the step is CUDA-graph capturable;
* MPIs are packed ``[B,S,H,W,4]``; source pass, target warp+composite and the losses are fused
  sm_100a kernels (``mine_b200.ops.cuda``) on GPU and the PyTorch spec on CPU;
* data parallelism = flat gradient arena + bucketed mean all-reduce overlapped with backward
  (own NVLink kernels or NCCL) + cross-replica BN statistics through the same communicator;
  static graph: no unused parameters, no buffer broadcast;
* exact resume (step, epoch, scheduler, RNG); evaluation sharded over the ranks (meters reduced on a host-side group).
"""
from __future__ import annotations

import glob
import itertools
import os
from typing import Dict, List, Mapping, Optional

import torch

from . import geometry as geo
from .config import get as cfg_get
from .models import checkpoint as ckpt
from .models.decoder import DepthDecoder
from .models.encoder import ResnetEncoder
from .models.norm import BatchNorm, flush_batch_counters, set_stat_reducer
from .ops import api as ops
from .optim import ArenaAdam, MultiStepLR
from .parallel import bootstrap
from .parallel.comm import Communicator, make_communicator
from .parallel.grad_sync import FlatArena, GradSync
from .spec import losses as L
from .spec import sampling as S
from .spec.embedder import get_embedder
from .utils import AverageMeter, NullLogger, PhaseProfiler, disparity_normalization_vis, run_shell_cmd
from .utils.misc import StopRequest
from .utils.misc import rng_state, set_rng_state

NO_SCALE_DATASETS = ("flowers", "kitti_raw", "dtu")     # disp_lambda = 0 and scale factor = 1

_LOSS_KEYS_TRAIN = ("loss", "loss_rgb_src", "loss_ssim_src", "loss_disp_pt3dsrc", "loss_rgb_tgt",
                    "loss_ssim_tgt", "lpips_tgt", "psnr_tgt", "loss_disp_pt3dtgt")
_LOSS_KEYS_VAL = _LOSS_KEYS_TRAIN[1:]


class PixelGrid:
    """Holds the ``3xHxW`` homogeneous pixel grid of one pyramid level.  External code reads
    ``task.homography_sampler_list[scale].meshgrid`` (``image_to_video.py:140``)."""

    def __init__(self, h: int, w: int, device=None):
        self.Height_tgt, self.Width_tgt = h, w
        self.meshgrid = geo.pixel_grid(h, w, device=device)


def bg_depth_inf(config: Mapping) -> bool:
    """Background-at-infinity switch with the reference's effective semantics (``synthesis_task.py:265,273,427,466``
    read ``mpi.render_tgt_rgb_depth``, which no upstream config defines): ``mpi.is_bg_depth_inf`` only counts when
    ``engine.honor_bg_depth_inf`` is set."""
    if bool(cfg_get(config, "mpi.render_tgt_rgb_depth", False)):
        return True
    return bool(cfg_get(config, "engine.honor_bg_depth_inf", False)) and bool(cfg_get(config, "mpi.is_bg_depth_inf", False))


def _get_disparity_list(config: Mapping, B: int, device=None, generator=None) -> torch.Tensor:
    return S.disparity_planes(config, B, device=device, generator=generator)


class _LossTerms:
    """Collects the differentiable scalar terms of all pyramid scales and assembles them with THREE small kernels
    (stack, ``w * t + b``, dot) instead of one framework op per addition / lambda multiplication (~75 scalar launches per
    step forward, ~25 backward).  ``add`` returns the position of the term; ``finish`` returns the weighted values (what
    the loss dictionary reports) and their selected sum (the training loss)."""
    _CONST: Dict = {}

    def __init__(self):
        self.raw, self.w, self.b, self.sel = [], [], [], []

    def add(self, raw: torch.Tensor, weight: float = 1.0, offset: float = 0.0, in_total: bool = True) -> int:
        self.raw.append(raw.reshape(()))
        self.w.append(float(weight)); self.b.append(float(offset)); self.sel.append(1.0 if in_total else 0.0)
        return len(self.raw) - 1

    @classmethod
    def _const(cls, values, device) -> torch.Tensor:
        key = (tuple(values), str(device))
        t = cls._CONST.get(key)
        if t is None:                      # first (eager) step: host -> device copies are illegal under graph capture
            t = cls._CONST[key] = torch.tensor(values, dtype=torch.float32, device=device)
        return t

    def finish(self):
        dev = self.raw[0].device
        stacked = torch.stack([r.float() for r in self.raw])
        vals = torch.addcmul(self._const(self.b, dev), stacked, self._const(self.w, dev))
        return vals, torch.dot(vals, self._const(self.sel, dev))


class SynthesisTask:
    def __init__(self, config: Dict, logger=None, is_val: bool = False, device=None,
                 comm: Optional[Communicator] = None):
        self.config = config
        self.logger = logger if logger is not None else NullLogger()
        self.is_val = is_val
        self.tb_writer = config.get("tb_writer", None)
        config.setdefault("global_rank", bootstrap.rank())
        if device is None:
            device = config.get("device", None)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)

        multires = int(config["model.pos_encoding_multires"])
        self.embedder, out_dim = get_embedder(multires)
        self.backbone = ResnetEncoder(num_layers=50, pretrained=bool(config.get("model.imagenet_pretrained", True))
                                      and not config.get("training.pretrained_checkpoint_path")).to(self.device)
        self.decoder = DepthDecoder(num_ch_enc=self.backbone.num_ch_enc, use_alpha=bool(config.get("mpi.use_alpha", False)),
                                    num_output_channels=4, scales=range(4), use_skips=True,
                                    embedder=None, embedder_out_dim=out_dim, multires=multires).to(self.device)
        # ImageNet initialisation: the reference downloads the weights, here they must be present locally.  A silent
        # random-init fallback changes convergence, so it is logged (training.log), recorded in the config that is
        # written next to the checkpoints, and fatal with model.require_imagenet_weights.
        if self.backbone.pretrained_requested and self.backbone.pretrained_source is None:
            msg = ("model.imagenet_pretrained=true but no local ResNet-50 weights were found (set MINE_RESNET50_WEIGHTS or "
                   "model.imagenet_pretrained=false): the encoder starts from RANDOM initialisation")
            if bool(config.get("model.require_imagenet_weights", False)):
                raise FileNotFoundError(msg)
            self.logger.warning(msg)
        config["model.imagenet_weights_used"] = (self.backbone.pretrained_source or "none (random initialisation)") \
            if self.backbone.pretrained_requested else "not requested"

        self.comm = comm if comm is not None else (
            Communicator() if is_val else make_communicator(cfg_get(config, "engine.comm", "auto"), self.device))
        # flat arenas: parameters, gradients (on the communicator's symmetric heap when it has one), moments
        n_b, n_d = len(list(self.backbone.parameters())), len(list(self.decoder.parameters()))
        self.arena = FlatArena(list(self.backbone.parameters()) + list(self.decoder.parameters()),
                               grad_alloc=getattr(self.comm, "alloc_symmetric", None))
        self.optimizer = ArenaAdam(self.arena, [n_b, n_d], [config["lr.backbone_lr"], config["lr.decoder_lr"]],
                                   weight_decay=float(config["lr.weight_decay"]))

        # restore on rank 0, then broadcast weights AND optimizer state (the reference restores the
        # optimizer on rank 0 only and lets replicas diverge)
        self.resume_meta: Dict = {}
        path = config.get("training.pretrained_checkpoint_path")
        if path and str(path).startswith("hdfs"):
            if config["global_rank"] == 0:
                run_shell_cmd(["hdfs", "dfs", "-get", path, "."], self.logger)
            path = os.path.basename(path)
            config["training.pretrained_checkpoint_path"] = path
        if path and config["global_rank"] == 0:
            self.resume_meta = ckpt.restore_model(path, self.backbone, self.decoder, self.optimizer, logger=self.logger)
        elif not path:
            self.logger.info("Not using pre-trained model...")
        if bootstrap.world_size() > 1 and not is_val:
            bootstrap.broadcast_module_state(self.backbone, self.decoder)
            if path:
                sd = bootstrap.broadcast_object(self.optimizer.state_dict() if config["global_rank"] == 0 else None)
                meta = bootstrap.broadcast_object(self.resume_meta if config["global_rank"] == 0 else None)
                if config["global_rank"] != 0:
                    self.optimizer.load_state_dict(sd)
                    self.resume_meta = meta

        if not is_val:
            set_stat_reducer(self.backbone, self.comm.allreduce_sum_ if self.comm.world_size > 1 else None)
            set_stat_reducer(self.decoder, self.comm.allreduce_sum_ if self.comm.world_size > 1 else None)
            self.grad_sync = GradSync(self.arena, self.comm)
            self.lr_scheduler = MultiStepLR(self.optimizer, config["lr.decay_steps"], gamma=float(config["lr.decay_gamma"]))
            self.backbone.train()
            self.decoder.train()
        else:
            self.grad_sync = None
            self.lr_scheduler = None
            self.backbone.eval()
            self.decoder.eval()

        from .engine import ModelRunner
        self.runner = ModelRunner(self.backbone, self.decoder, config, self.device)

        H, W = int(config["data.img_h"]), int(config["data.img_w"])
        self.homography_sampler_list = [PixelGrid(H // 2 ** s, W // 2 ** s, self.device) for s in range(4)]
        self.upsample_list = [(lambda x, s=s: L.nearest_downsample(x, s)) for s in range(4)]
        self.ssim = L.SSIM(size_average=True)
        self.lpips = None
        if config["global_rank"] == 0:
            from .models.lpips import build_lpips
            self.lpips = build_lpips(self.device, self.logger)

        self.init_data(self.device)
        self.train_losses = {k: AverageMeter("train_" + k) for k in _LOSS_KEYS_TRAIN}
        self.val_losses = {k: AverageMeter("val_" + k) for k in _LOSS_KEYS_VAL}
        self.current_epoch = 0
        self.epoch_step = 0                     # batches of the current epoch already consumed (mid-epoch resume)
        self.stop_request = None                # utils.misc.StopRequest installed by train(): SIGTERM -> save + exit
        self.stopped = False
        self.global_step = 0
        self.profiler = PhaseProfiler(enabled=False)
        self._gen = None
        self._graph, self._graph_out, self._static = None, None, None
        self._want_graph = bool(cfg_get(config, "engine.cuda_graph", False)) and not is_val
        # Resume (step / epoch / scheduler / RNG) only from a full training state, i.e. a checkpoint that also holds the
        # optimizer (``checkpoint_latest.pth``).  A weights-only ``training.pretrained_checkpoint_path`` is a
        # fine-tuning start exactly as upstream: weights loaded, counters and learning rate fresh.
        if self.resume_meta and self.resume_meta.get("has_optimizer", False) and cfg_get(config, "engine.resume", True):
            self._apply_resume(self.resume_meta)
        elif self.resume_meta:
            self.resume_meta = {}

    # ------------------------------------------------------------------------------------------
    # data
    # ------------------------------------------------------------------------------------------
    def init_data(self, device):
        c = self.config
        B, H, W = int(c["data.per_gpu_batch_size"]), int(c["data.img_h"]), int(c["data.img_w"])
        Lv, N = int(c["data.num_tgt_views"]), int(c["data.visible_point_count"])
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self.src_imgs, self.K_src, self.K_src_inv, self.pt3d_src = z(B, 3, H, W), z(B, 3, 3), z(B, 3, 3), z(B, 3, N)
        self.tgt_imgs, self.G_src_tgt = z(B, Lv, 3, H, W), z(B, Lv, 4, 4)
        self.K_tgt, self.K_tgt_inv, self.pt3d_tgt = z(B, Lv, 3, 3), z(B, Lv, 3, 3), z(B, Lv, 3, N)
        self.G_tgt_src = z(B, 4, 4)

    def set_data(self, items):
        """Accepts host (pageable or pinned) or device tensors in the reference batch format.
        Copies are asynchronous; nothing here synchronises the host."""
        src, tgt = items
        d = self.device
        mv = lambda t: t.to(d, dtype=torch.float32, non_blocking=True)
        self.src_imgs, self.K_src, self.K_src_inv = mv(src["img"]), mv(src["K"]), mv(src["K_inv"])
        self.pt3d_src = mv(src["xyzs"])
        tgt_imgs, g_src_tgt = mv(tgt["img"]), mv(tgt["G_src_tgt"])
        if tgt_imgs.shape[1] != 1:
            raise ValueError("exactly one target view per source is supported (data.num_tgt_views = 1)")
        self.tgt_imgs, self.G_src_tgt = tgt_imgs[:, 0], g_src_tgt[:, 0]
        self.K_tgt, self.K_tgt_inv, self.pt3d_tgt = mv(tgt["K"])[:, 0], mv(tgt["K_inv"])[:, 0], mv(tgt["xyzs"])[:, 0]
        self.G_tgt_src = geo.inv_affine4x4(self.G_src_tgt)
        # per-level intrinsics / image pyramids once per batch (not once per scale inside the loss)
        self._K_src_lv = geo.intrinsics_pyramid(self.K_src)
        self._K_tgt_lv = geo.intrinsics_pyramid(self.K_tgt)
        self._K_src_inv_lv = geo.inv3x3(self._K_src_lv)
        self._src_pyr = [L.nearest_downsample(self.src_imgs, s) for s in range(4)]
        self._tgt_pyr = [L.nearest_downsample(self.tgt_imgs, s) for s in range(4)]

    # ------------------------------------------------------------------------------------------
    # model
    # ------------------------------------------------------------------------------------------
    def mpi_predictor(self, src_imgs_BCHW: torch.Tensor, disparity_BS: torch.Tensor) -> List[torch.Tensor]:
        """-> 4 MPIs ``[B,S,4,H/2^s,W/2^s]`` (zero-copy views of the packed tensors)."""
        return [ops.unpack_mpi(p) for p in self.runner.predict(src_imgs_BCHW, disparity_BS)]

    def network_forward(self) -> Dict:
        c = self.config
        B = self.src_imgs.shape[0]
        disparity = _get_disparity_list(c, B, device=self.src_imgs.device, generator=self._gen)
        s_fine = int(c["mpi.num_bins_fine"])
        if s_fine > 0:
            with torch.no_grad():
                coarse = self.runner.predict(self.src_imgs, disparity)[0]
                w = ops.plane_weights_mean(coarse, disparity, self.K_src_inv, self._bg_inf())
                w5 = w[:, :, None, None, None]
                disparity = S.refine_disparity(disparity, w5, s_fine, generator=self._gen)
        mpis = self.mpi_predictor(self.src_imgs, disparity)
        return {"mpi_all_src_list": mpis, "disparity_all_src": disparity}

    def _bg_inf(self) -> bool:
        return bg_depth_inf(self.config)

    # ------------------------------------------------------------------------------------------
    # rendering
    # ------------------------------------------------------------------------------------------
    def render_novel_view(self, mpi_all_rgb_src, mpi_all_sigma_src, disparity_all_src, G_tgt_src,
                          K_src_inv, K_tgt, scale=0, scale_factor=None) -> Dict[str, torch.Tensor]:
        if scale_factor is not None:
            if not torch.is_tensor(scale_factor):
                scale_factor = torch.full((G_tgt_src.shape[0],), float(scale_factor), device=G_tgt_src.device)
            G_tgt_src = geo.rescale_translation(G_tgt_src, scale_factor)
        packed = ops.pack_rgb_sigma(mpi_all_rgb_src, mpi_all_sigma_src)
        rgb, depth, mask = ops.render_tgt(packed, disparity_all_src, G_tgt_src, K_src_inv, K_tgt,
                                          bool(self.config.get("mpi.use_alpha", False)), self._bg_inf())
        return {"tgt_imgs_syn": rgb, "tgt_disparity_syn": torch.reciprocal(depth), "tgt_mask_syn": mask}

    def _zero_scalar(self, device) -> torch.Tensor:
        """A cached fp32 zero (reported for the loss terms whose lambda is 0: no fill kernel per scale and step)."""
        z = getattr(self, "_zero_cache", None)
        if z is None or z.device != device:
            z = self._zero_cache = torch.zeros((), device=device)
        return z

    def compute_scale_factor(self, disparity_syn_pt3dsrc, pt3d_disp_src):
        if self.config["data.name"] in NO_SCALE_DATASETS:
            return torch.ones(pt3d_disp_src.shape[0], dtype=torch.float32, device=pt3d_disp_src.device)
        return L.scale_factor_from_points(disparity_syn_pt3dsrc, pt3d_disp_src)

    # ------------------------------------------------------------------------------------------
    # losses
    # ------------------------------------------------------------------------------------------
    def loss_fcn_per_scale(self, scale, mpi_all_src, disparity_all_src, scale_factor=None, is_val=False, _terms=None):
        """Losses of one pyramid scale.  ``_terms`` (internal, set by :meth:`loss_fcn`): the differentiable terms are
        recorded raw in a :class:`_LossTerms` collector - the dictionary then maps their names to collector positions and
        carries no ``"loss"`` entry; :meth:`loss_fcn` assembles everything for all scales at once."""
        c = self.config
        src_img, tgt_img = self._src_pyr[scale], self._tgt_pyr[scale]
        B = src_img.shape[0]
        K_src, K_tgt, K_src_inv = self._K_src_lv[scale], self._K_tgt_lv[scale], self._K_src_inv_lv[scale]
        use_alpha = bool(c.get("mpi.use_alpha", False))
        if tuple(mpi_all_src.shape[-2:]) != tuple(src_img.shape[-2:]):
            raise ValueError("MPI resolution must equal the image resolution of its pyramid level")

        # --- source view: composite, source blending, disparity --------------------------------
        with self.profiler.phase("render_src"):
            src = ops.render_src(ops.pack_mpi(mpi_all_src), disparity_all_src, K_src_inv, src_img, use_alpha,
                                 self._bg_inf(), blend=bool(c.get("training.src_rgb_blending", True)))
        src_imgs_syn, src_disparity_syn, mpi_for_tgt = src["rgb"], src["disparity"], src["mpi"]

        # --- sparse points in the source frame, scale calibration ------------------------------
        loss_disp_src, scale_factor = ops.sparse_point_loss(src_disparity_syn, K_src, self.pt3d_src, scale_factor,
                                                            calibrate=c["data.name"] not in NO_SCALE_DATASETS)

        # --- target view (true dependency on the scale factor) --------------------
        with self.profiler.phase("render_tgt"):
            packed = mpi_for_tgt
            g = geo.rescale_translation(self.G_tgt_src, scale_factor)
            tgt_rgb, tgt_depth, tgt_mask = ops.render_tgt(packed, disparity_all_src, g, K_src_inv, K_tgt,
                                                          use_alpha, self._bg_inf())
        tgt_disparity_syn = torch.reciprocal(tgt_depth)

        # --- losses -----------------------------------------------------------------------
        with self.profiler.phase("losses"):
            disp_lambda = 0.0 if c["data.name"] in NO_SCALE_DATASETS else 1.0
            lam1 = float(c.get("loss.smoothness_lambda_v1", 0.5))
            lam2 = float(c.get("loss.smoothness_lambda_v2", 1.0))
            gmin, ratio = float(c["loss.smoothness_gmin"]), float(c.get("loss.smoothness_grad_ratio", 0.1))
            with torch.no_grad():
                loss_rgb_src = (src_imgs_syn - src_img).abs().mean()
                loss_ssim_src = 1 - ops.ssim(src_imgs_syn, src_img)
                loss_smooth_src = ops.edge_aware_loss(src_img, src_disparity_syn, gmin, ratio)
            raw_disp_tgt = ops.sparse_point_loss(tgt_disparity_syn, K_tgt, self.pt3d_tgt, scale_factor)[0]
            raw_rgb_tgt = ops.masked_l1(tgt_rgb, tgt_img, tgt_mask, float(c["mpi.valid_mask_threshold"]))
            raw_smooth_tgt = ops.edge_aware_loss(tgt_img, tgt_disparity_syn, gmin, ratio) if lam1 != 0.0 else None
            raw_smooth_tgt_v2 = ops.edge_aware_loss_v2(tgt_img, tgt_disparity_syn) if lam2 != 0.0 else None
            raw_smooth_src_v2 = ops.edge_aware_loss_v2(src_img, src_disparity_syn) if lam2 != 0.0 else None
            raw_ssim_tgt = ops.ssim(tgt_rgb, tgt_img)
            if _terms is not None:
                # cross-scale rule of the reference: scale 0 counts every term; scales 1..3 count the sparse-disparity and
                # v2-smoothness terms, plus RGB / SSIM under training.use_multi_scale, and never the v1 smoothness
                multi = scale == 0 or bool(c.get("training.use_multi_scale", True))
                zero = self._zero_scalar(tgt_rgb.device)
                pos = {"loss_disp_pt3dtgt": _terms.add(raw_disp_tgt, disp_lambda),
                       "loss_disp_pt3dsrc": _terms.add(loss_disp_src, disp_lambda),
                       "loss_rgb_tgt": _terms.add(raw_rgb_tgt, 1.0, 0.0, multi),
                       "loss_ssim_tgt": _terms.add(raw_ssim_tgt, -1.0, 1.0, multi)}
                if raw_smooth_tgt is not None:
                    pos["loss_smooth_tgt"] = _terms.add(raw_smooth_tgt, lam1, 0.0, scale == 0)
                if raw_smooth_tgt_v2 is not None:
                    pos["loss_smooth_tgt_v2"] = _terms.add(raw_smooth_tgt_v2, lam2)
                    pos["loss_smooth_src_v2"] = _terms.add(raw_smooth_src_v2, lam2)
                with torch.no_grad():
                    if is_val and scale == 0 and self.lpips is not None:
                        lpips_tgt = self.lpips(tgt_rgb, tgt_img).mean()
                    else:
                        lpips_tgt = zero
                    psnr_tgt = ops.psnr(tgt_rgb, tgt_img)
                loss_dict = {"_pos": pos, "loss_rgb_src": loss_rgb_src, "loss_ssim_src": loss_ssim_src,
                             "loss_smooth_src": loss_smooth_src, "loss_smooth_tgt": zero, "loss_smooth_src_v2": zero,
                             "loss_smooth_tgt_v2": zero, "lpips_tgt": lpips_tgt, "psnr_tgt": psnr_tgt}
                vis = {"src_disparity_syn": src_disparity_syn, "tgt_disparity_syn": tgt_disparity_syn,
                       "tgt_imgs_syn": tgt_rgb, "tgt_mask_syn": tgt_mask, "src_imgs_syn": src_imgs_syn}
                return loss_dict, vis, scale_factor
            loss_disp_src = disp_lambda * loss_disp_src
            loss_disp_tgt = disp_lambda * raw_disp_tgt
            loss_rgb_tgt = raw_rgb_tgt
            zero = torch.zeros((), device=tgt_rgb.device)
            loss_smooth_tgt = lam1 * raw_smooth_tgt if raw_smooth_tgt is not None else zero
            loss_smooth_tgt_v2 = lam2 * raw_smooth_tgt_v2 if raw_smooth_tgt_v2 is not None else zero
            loss_smooth_src_v2 = lam2 * raw_smooth_src_v2 if raw_smooth_src_v2 is not None else zero
            loss_ssim_tgt = 1 - raw_ssim_tgt
            with torch.no_grad():
                if is_val and scale == 0 and self.lpips is not None:
                    lpips_tgt = self.lpips(tgt_rgb, tgt_img).mean()
                else:
                    lpips_tgt = torch.zeros((), device=tgt_rgb.device)
                psnr_tgt = ops.psnr(tgt_rgb, tgt_img)
            loss = (loss_disp_tgt + loss_disp_src + loss_rgb_tgt + loss_ssim_tgt + loss_smooth_tgt
                    + loss_smooth_src_v2 + loss_smooth_tgt_v2)

        loss_dict = {"loss": loss, "loss_rgb_src": loss_rgb_src, "loss_ssim_src": loss_ssim_src,
                     "loss_disp_pt3dsrc": loss_disp_src, "loss_smooth_src": loss_smooth_src,
                     "loss_smooth_tgt": loss_smooth_tgt, "loss_smooth_src_v2": loss_smooth_src_v2,
                     "loss_smooth_tgt_v2": loss_smooth_tgt_v2, "loss_rgb_tgt": loss_rgb_tgt,
                     "loss_ssim_tgt": loss_ssim_tgt, "lpips_tgt": lpips_tgt, "psnr_tgt": psnr_tgt,
                     "loss_disp_pt3dtgt": loss_disp_tgt}
        vis = {"src_disparity_syn": src_disparity_syn, "tgt_disparity_syn": tgt_disparity_syn,
               "tgt_imgs_syn": tgt_rgb, "tgt_mask_syn": tgt_mask, "src_imgs_syn": src_imgs_syn}
        return loss_dict, vis, scale_factor

    def loss_fcn(self, is_val: bool):
        with self.profiler.phase("network"):
            endpoints = self.network_forward()
        per_scale, vis_list, scale_factor = [], [], None
        fused = os.environ.get("MINE_B200_LOSS_ASSEMBLY", "fused") == "fused"
        terms = _LossTerms() if fused else None
        for scale in range(4):
            ld, vis, scale_factor = self.loss_fcn_per_scale(scale, endpoints["mpi_all_src_list"][scale],
                                                            endpoints["disparity_all_src"], scale_factor, is_val=is_val,
                                                            _terms=terms)
            per_scale.append(ld)
            vis_list.append(vis)
        if fused:
            vals, total = terms.finish()                   # all scalar arithmetic of the step: three small kernels
            unbound = vals.unbind(0)
            loss_dict = per_scale[0]
            for name, i in loss_dict.pop("_pos").items():
                loss_dict[name] = unbound[i]
            loss_dict["loss"] = total
            return loss_dict, vis_list[0]
        loss_dict = per_scale[0]
        total = loss_dict["loss"]
        for s in range(1, 4):
            if self.config.get("training.use_multi_scale", True):
                total = total + per_scale[s]["loss_rgb_tgt"] + per_scale[s]["loss_ssim_tgt"]
            total = total + per_scale[s]["loss_disp_pt3dsrc"] + per_scale[s]["loss_disp_pt3dtgt"]
            total = total + per_scale[s]["loss_smooth_src_v2"] + per_scale[s]["loss_smooth_tgt_v2"]
        loss_dict["loss"] = total
        return loss_dict, vis_list[0]

    # ------------------------------------------------------------------------------------------
    # training
    # ------------------------------------------------------------------------------------------
    def train_step(self, items) -> Dict[str, torch.Tensor]:
        """One optimisation step on a batch in the reference format (no host sync).  With
        ``engine.cuda_graph`` the whole step (forward, losses, backward, Adam) is ONE graph replay."""
        graph_ok = self.comm.world_size == 1 or getattr(self.comm, "graph_safe", False)
        if self._graph is None and self._want_graph and graph_ok and self.device.type == "cuda":
            self.enable_cuda_graph(items)
        if self._graph is not None:
            self._load_static(items)
            self._graph.replay()
            self.global_step += 1
            self.optimizer.step_count += 1
            return self._graph_out
        return self._train_step_eager(items)

    # -- CUDA graph capture of the full step -------------------------------------------------------
    def _load_static(self, items) -> None:
        for dst, src in zip(self._static, items):
            for k, t in dst.items():
                t.copy_(src[k], non_blocking=True)

    def enable_cuda_graph(self, example_items, warmup: int = 3) -> None:
        """Capture ``_train_step_eager`` (launch-bound: ~1.5 k kernels of a few microseconds each) into a
        CUDA graph.  Model / optimizer / BN state is snapshotted around the warm-up replays so enabling the
        graph does not consume training steps.  Host-side scalars that change per step live on the device
        (Adam's step counter / lr, BN counters), so replays stay exact."""
        dev = self.device
        keys_src = ("img", "K", "K_inv", "xyzs")
        keys_tgt = ("img", "K", "K_inv", "xyzs", "G_src_tgt")
        src, tgt = example_items
        self._static = ({k: src[k].to(dev, dtype=torch.float32).clone() for k in keys_src},
                        {k: tgt[k].to(dev, dtype=torch.float32).clone() for k in keys_tgt})
        bufs = [self.arena.data, self.optimizer.exp_avg, self.optimizer.exp_avg_sq]
        bufs += [b for m in (self.backbone, self.decoder) for b in m.buffers()]
        self.optimizer.sync_hyper()
        bufs += list(self.optimizer._hyper)
        snap = [b.clone() for b in bufs]
        step0, gstep0 = self.optimizer.step_count, self.global_step
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._train_step_eager(self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from .ops import cuda as _K
        n0 = _K.LAUNCHES["count"]
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._train_step_eager(self._static)
        self.launches_per_step = _K.LAUNCHES["count"] - n0        # our kernels inside one replay
        with torch.no_grad():
            for b, s_ in zip(bufs, snap):
                b.copy_(s_)
        self.optimizer.step_count, self.global_step = step0, gstep0
        self._graph, self._graph_out = graph, out

    def _train_step_eager(self, items) -> Dict[str, torch.Tensor]:
        self.global_step += 1
        self.set_data(items)
        self.grad_sync.begin_step()
        self.optimizer.zero_grad()
        from .ops import conv_engine as _E
        BatchNorm.defer_counters = True
        _E.defer_running_stats(True)               # running-statistic updates of all layers: one launch after the forward
        try:
            loss_dict, _ = self.loss_fcn(is_val=False)
        finally:
            BatchNorm.defer_counters = False
            _E.defer_running_stats(False)          # flushes
        flush_batch_counters(self.backbone, self.decoder)
        with self.profiler.phase("backward"):
            loss_dict["loss"].backward()
        with self.profiler.phase("grad_sync"):
            self.grad_sync.finish()
        with self.profiler.phase("optimizer"):
            self.optimizer.step()
        return loss_dict

    def _is_main(self) -> bool:
        return self.config.get("global_rank", 0) == 0

    def train_epoch(self, train_data_loader, val_data_loader, epoch):
        sampler = getattr(train_data_loader, "sampler", None)
        if sampler is not None and hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        self.backbone.train()
        self.decoder.train()
        self.current_epoch = epoch
        self.config["current_epoch"] = epoch
        for m in self.train_losses.values():
            m.reset()
        c = self.config
        log_every = int(cfg_get(c, "training.log_interval", 10))
        ckpt_every = int(cfg_get(c, "training.checkpoint_interval", 5000))
        max_steps = int(cfg_get(c, "training.max_steps", 0))
        # resuming inside an epoch: this rank's shard continues after the batches it had consumed
        skip = self.epoch_step if self.epoch_step < len(train_data_loader) else 0
        if skip:
            if hasattr(sampler, "set_start"):
                sampler.set_start(skip * int(c["data.per_gpu_batch_size"]))
                batches = train_data_loader
            else:
                batches = itertools.islice(iter(train_data_loader), skip, None)
        else:
            batches = train_data_loader
        self.epoch_step = skip
        for step, items in enumerate(batches, start=skip + 1):
            loss_dict = self.train_step(items)
            self.epoch_step = step
            if step % log_every == 0 and self._is_main():
                self.log_training(epoch, step, self.global_step, len(train_data_loader), loss_dict)
            if step % ckpt_every == 0 and self._is_main():
                self.save_checkpoint("checkpoint_latest.pth", with_optimizer=True)
            if self.global_step > 0 and (self.global_step == 2000
                                         or self.global_step % int(c["training.eval_interval"]) == 0):
                self._eval_and_checkpoint(val_data_loader)
            if max_steps and self.global_step >= max_steps:
                return self.epoch_step >= len(train_data_loader)
            if step % log_every == 0 and self.stop_request is not None:
                # all ranks reach this poll at the same step and leave together (no rank is left in a collective)
                if bootstrap.any_rank(self.stop_request.is_set(), self.device):
                    self.stopped = True
                    return self.epoch_step >= len(train_data_loader)
        return True

    def _eval_and_checkpoint(self, val_data_loader):
        """All ranks arrive here at the same step.  With ``training.all_rank_eval`` (default) and more than one rank the
        validation batches are dealt round-robin to the ranks and the meters are summed over ranks (BN in eval mode does not
        communicate, so the ranks evaluate independently); otherwise rank 0 evaluates alone.  Either way every rank then
        waits at the host-side barrier - no rank runs ahead into a collective (the reference relies on accidental
        pairing, SURVEY 2.4 'rank-asymmetric control flow')."""
        has_val = val_data_loader is not None and len(val_data_loader) > 0
        world = bootstrap.world_size()
        sharded = has_val and world > 1 and bool(self.config.get("training.all_rank_eval", True))
        if sharded:
            self.run_eval(val_data_loader, shard=(bootstrap.rank(), world))
        if self._is_main() and has_val:
            if not sharded:
                self.run_eval(val_data_loader)
            path = self.save_checkpoint("checkpoint_%012d.pth" % self.global_step, with_optimizer=False)
            if "hdfs_workspace" in self.config and path:
                tb = sorted(glob.glob(os.path.join(self.config["local_workspace"], "events.out.tfevents.*")))
                for f in [path, self.config.get("log_file")] + tb[-1:]:
                    if f:
                        run_shell_cmd(["hdfs", "dfs", "-put", "-f", f, self.config["hdfs_workspace"]], self.logger)
        bootstrap.patient_barrier()                # host-side group with a 24 h timeout: validation may take long
        bootstrap.barrier()                        # ... then re-align the device streams of all ranks

    def save_checkpoint(self, name: str, with_optimizer: bool) -> Optional[str]:
        ws = self.config.get("local_workspace")
        if not ws:
            return None
        path = os.path.join(ws, name)
        meta = {"global_step": self.global_step, "epoch": self.current_epoch, "epoch_step": self.epoch_step,
                "lr_scheduler": self.lr_scheduler.state_dict() if self.lr_scheduler else None,
                "rng": rng_state() if with_optimizer else None}
        ckpt.save_checkpoint(path, self.backbone, self.decoder, self.optimizer if with_optimizer else None, meta)
        self.logger.info("Checkpoint saved at {}".format(path))
        if with_optimizer and "hdfs_workspace" in self.config:
            for f in (path, self.config.get("log_file")):
                if f:
                    run_shell_cmd(["hdfs", "dfs", "-put", "-f", f, self.config["hdfs_workspace"]], self.logger)
        return path

    def _apply_resume(self, meta: Mapping) -> None:
        self.global_step = int(meta.get("global_step", 0))
        self.current_epoch = int(meta.get("epoch", 0))
        self.epoch_step = int(meta.get("epoch_step", 0))
        if self.lr_scheduler is not None and meta.get("lr_scheduler"):
            self.lr_scheduler.load_state_dict(meta["lr_scheduler"])
        # the saved generator state is rank 0's: restoring it everywhere would make all replicas draw identical
        # disparity samples from then on, so the other ranks keep their own (seeded per rank) streams
        if meta.get("rng") is not None and int(self.config.get("global_rank", 0)) == 0:
            try:
                set_rng_state(meta["rng"])
            except Exception:      # RNG layouts differ across devices/torch versions: not fatal
                pass
        self.logger.info("Resumed at epoch %d (batch %d), global_step %d" % (self.current_epoch, self.epoch_step,
                                                                              self.global_step))

    def train(self, train_data_loader, val_data_loader):
        max_steps = int(cfg_get(self.config, "training.max_steps", 0))
        start_epoch = max(self.current_epoch, 1) if self.resume_meta else 1
        if self.stop_request is None:
            self.stop_request = StopRequest()
        for epoch in range(start_epoch, int(self.config["training.epochs"]) + 1):
            finished = self.train_epoch(train_data_loader, val_data_loader, epoch)
            if finished:
                self.lr_scheduler.step()
                self.current_epoch, self.epoch_step = epoch + 1, 0        # a restart begins the next epoch
            if self._is_main():
                self.logger.info("Epoch finished, average losses: " if finished else
                                 ("Stop requested (signal): state saved, exiting. " if self.stopped
                                  else "Stopped at training.max_steps: "))
                for v in self.train_losses.values():
                    self.logger.info("    {}".format(v))
                # the state a restart continues from (upstream only writes it every checkpoint interval)
                self.save_checkpoint("checkpoint_latest.pth", with_optimizer=True)
            if self.stopped or (max_steps and self.global_step >= max_steps):
                break
        self.stop_request.uninstall()

    # ------------------------------------------------------------------------------------------
    # evaluation / logging
    # ------------------------------------------------------------------------------------------
    def run_eval(self, val_data_loader, shard=None):
        """Validation pass.  ``shard = (rank, world)``: this rank evaluates batches ``rank, rank + world, ...`` and the
        meter sums / counts are added up over all ranks at the end (every rank then holds the global averages)."""
        self.logger.info("Start running evaluation on validation set:")
        self.backbone.eval()
        self.decoder.eval()
        for m in self.val_losses.values():
            m.reset()
        with torch.no_grad():
            for step, items in enumerate(val_data_loader):
                if (step + 1) % 20 == 0:
                    self.logger.info("    Eval progress: {}/{}".format(step + 1, len(val_data_loader)))
                if shard is not None and step % shard[1] != shard[0]:
                    continue
                self.set_data(items)
                loss_dict, vis = self.loss_fcn(is_val=True)
                self.log_val(step, loss_dict, vis)
            if shard is not None:
                meters = list(self.val_losses.values())
                tot = bootstrap.allreduce_host_sum([m.sum for m in meters] + [float(m.count) for m in meters])
                for i, m in enumerate(meters):
                    m.sum, m.count = tot[i], int(round(tot[len(meters) + i]))
                    m.avg = m.sum / max(m.count, 1)
            self.logger.info("Evaluation finished, average losses: ")
            for v in self.val_losses.values():
                self.logger.info("    {}".format(v))
            if self.tb_writer is not None:
                for k, v in self.val_losses.items():
                    self.tb_writer.add_scalar(k + "/val", v.avg, self.global_step)
        self.backbone.train()
        self.decoder.train()

    def log_val(self, step, loss_dict, visualization_dict, max_image_batches: int = 4):
        B = self.src_imgs.shape[0]
        vals = torch.stack([loss_dict[k].detach().float().reshape(()) for k in self.val_losses]).cpu()   # one D2H
        for (k, m), v in zip(self.val_losses.items(), vals.tolist()):
            m.update(v, n=B)
        if self.tb_writer is None or step >= max_image_batches:
            return
        import torchvision
        grid = torchvision.utils.make_grid
        if self.global_step == int(self.config["training.eval_interval"]):
            self.tb_writer.add_image("00_src_images", grid(self.src_imgs), step)
            self.tb_writer.add_image("01_gt_tgt_images", grid(self.tgt_imgs), step)
        tag = "step_%d" % self.global_step
        self.tb_writer.add_image("02_syn_src_images/" + tag, grid(visualization_dict["src_imgs_syn"]), step)
        self.tb_writer.add_image("03_syn_src_disparity_map/" + tag,
                                 grid(disparity_normalization_vis(visualization_dict["src_disparity_syn"])), step)
        self.tb_writer.add_image("04_syn_tgt_images/" + tag, grid(visualization_dict["tgt_imgs_syn"]), step)
        self.tb_writer.add_image("05_syn_tgt_disparity_map/" + tag,
                                 grid(disparity_normalization_vis(visualization_dict["tgt_disparity_syn"])), step)

    def log_training(self, epoch, step, global_step, dataset_length, loss_dict):
        keys = ["loss", "loss_rgb_src", "loss_ssim_src", "loss_smooth_src", "loss_disp_pt3dsrc", "loss_rgb_tgt",
                "loss_ssim_tgt", "loss_smooth_tgt", "loss_disp_pt3dtgt", "lpips_tgt", "psnr_tgt"]
        vals = torch.stack([loss_dict[k].detach().float().reshape(()) for k in keys]).cpu().tolist()   # one D2H
        v = dict(zip(keys, vals))
        if not all(x == x and abs(x) != float("inf") for x in vals):
            # failure detection the reference lacks: stop at the first logged non-finite loss instead of
            # training on garbage; checkpoint_latest.pth still holds the last healthy state
            bad = [k for k, x in v.items() if not (x == x and abs(x) != float("inf"))]
            self.logger.info("non-finite training loss at global_step %d: %s" % (global_step, ", ".join(bad)))
            raise FloatingPointError("non-finite loss terms at global_step %d: %s" % (global_step, ", ".join(bad)))
        perf = self._throughput(global_step)
        self.logger.info(
            "epoch [%.3d] step [%d/%d] global_step = %d total_loss = %.4f encoder_lr = %.7f\n"
            "        src: rgb = %.4f\n        src: ssim = %.4f\n        src: smooth = %.4f\n"
            "        src: disp_pt3d = %.4f\n        tgt: rgb = %.4f\n        tgt: ssim = %.4f\n"
            "        tgt: smooth = %.4f\n        tgt: disp_pt3d = %.4f" %
            (epoch, step, dataset_length, global_step, v["loss"], self.optimizer.param_groups[0]["lr"],
             v["loss_rgb_src"], v["loss_ssim_src"], v["loss_smooth_src"], v["loss_disp_pt3dsrc"],
             v["loss_rgb_tgt"], v["loss_ssim_tgt"], v["loss_smooth_tgt"], v["loss_disp_pt3dtgt"]))
        if perf:
            self.logger.info("        perf: %.2f ms/step, %.1f images/s (this rank), peak memory %.2f GB" %
                             (perf["ms_per_step"], perf["images_per_s"], perf["peak_mem_gb"]))
        for k, m in self.train_losses.items():
            if self.tb_writer is not None:
                self.tb_writer.add_scalar(k + "/train", v[k], global_step)
            m.update(v[k])
        if self.tb_writer is not None:
            for k, x in perf.items():
                self.tb_writer.add_scalar("perf/" + k, x, global_step)

    def _throughput(self, global_step: int) -> Dict[str, float]:
        """Step time since the previous log line (CUDA events on the compute stream - the log line's own D2H is
        the only synchronisation), images/s of this rank and peak allocator memory (the reference
        reports no throughput / latency / memory figures at all)."""
        cuda = self.device.type == "cuda"
        import time
        now = torch.cuda.Event(enable_timing=True) if cuda else time.perf_counter()
        if cuda:
            now.record()
        out: Dict[str, float] = {}
        prev = getattr(self, "_perf_mark", None)
        if prev is not None and global_step > prev[1]:
            if cuda:
                now.synchronize()
                ms = prev[0].elapsed_time(now)
            else:
                ms = (now - prev[0]) * 1e3
            steps = global_step - prev[1]
            bsz = int(self.config["data.per_gpu_batch_size"])
            out = {"ms_per_step": ms / steps, "images_per_s": bsz * steps / max(ms, 1e-9) * 1e3,
                   "peak_mem_gb": torch.cuda.max_memory_allocated(self.device) / 2 ** 30 if cuda else 0.0}
        self._perf_mark = (now, global_step)
        return out
