"""Checkpoint adapter: reference-format ``{"backbone","decoder"[,"optimizer"]}`` <-> our modules.

Format facts (reference ``synthesis_task.py:629-631,650-652``, ``utils.py:40-67``):
keys may carry a ``module.`` prefix (saved from DDP wrappers); decoder ConvBlocks live in an
``nn.ModuleDict`` whose keys are ``'-'.join(str(("upconv", i, j)))`` - i.e. the *characters* of the
tuple's repr joined by dashes - e.g. ``convs.(-'-u-p-c-o-n-v-'-,- -4-,- -0-).conv.conv.weight``;
the backbone has ``encoder.fc.*`` entries that nothing uses.  We additionally store a ``"meta"``
entry (step / epoch / lr-scheduler / RNG) so that resume is exact - the reference resumes weights
only and ignores unknown top-level keys, so files stay loadable by it.
"""
from __future__ import annotations

import os
from typing import Dict, Mapping, Optional

import torch

FC_SHAPES = {"encoder.fc.weight": (1000, 2048), "encoder.fc.bias": (1000,)}
N_BACKBONE_PARAMS_REF = 161          # incl. fc.weight, fc.bias at positions 159, 160
N_BACKBONE_PARAMS = 159


def mangle(key_tuple) -> str:
    """The reference's accidental ModuleDict key: characters of ``str(tuple)`` joined by '-'."""
    return "-".join(str(key_tuple))


def _decoder_name_maps(decoder) -> Dict[str, str]:
    """our-name-prefix -> reference-name-prefix for every decoder sub-module."""
    m = {}
    for name in decoder.blocks.keys():
        _, i, j = name.split("_")
        m[f"blocks.{name}."] = f"convs.{mangle(('upconv', int(i), int(j)))}."
    for name in decoder.heads.keys():
        s = int(name.split("_")[1])
        m[f"heads.{name}."] = f"convs.{mangle(('dispconv', s))}."
    return m


def _rename(key: str, table: Mapping[str, str]) -> str:
    for a, b in table.items():
        if key.startswith(a):
            return b + key[len(a):]
    return key


def strip_module_prefix(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}


def backbone_to_reference(backbone, module_prefix: bool = True, include_fc: bool = True) -> Dict[str, torch.Tensor]:
    sd = {k: v.detach().contiguous().clone() for k, v in backbone.state_dict().items()}
    if include_fc:
        for k, shape in FC_SHAPES.items():
            sd[k] = torch.zeros(shape)
    return {("module." + k if module_prefix else k): v for k, v in sd.items()}


def decoder_to_reference(decoder, module_prefix: bool = True) -> Dict[str, torch.Tensor]:
    table = _decoder_name_maps(decoder)
    out = {}
    for k, v in decoder.state_dict().items():
        rk = _rename(k, table)
        out["module." + rk if module_prefix else rk] = v.detach().contiguous().clone()
    return out


def load_backbone(backbone, sd: Mapping[str, torch.Tensor]):
    sd = {k: v for k, v in strip_module_prefix(sd).items() if not k.startswith("encoder.fc.")}
    return backbone.load_state_dict(sd, strict=False)


def load_decoder(decoder, sd: Mapping[str, torch.Tensor]):
    inverse = {b: a for a, b in _decoder_name_maps(decoder).items()}
    sd = {_rename(k, inverse): v for k, v in strip_module_prefix(sd).items()}
    return decoder.load_state_dict(sd, strict=False)


# ---- optimizer state: torch.optim.Adam layout, param indices in reference order ---------------
def optimizer_to_reference(opt_sd: Mapping) -> Dict:
    """Insert zero state for the two ``fc`` params so indices match the reference's 161+60 layout."""
    state = {}
    for idx, st in opt_sd["state"].items():
        idx = int(idx)
        state[idx if idx < N_BACKBONE_PARAMS else idx + 2] = st
    any_step = next(iter(opt_sd["state"].values()), {}).get("step", torch.tensor(0.0)) if opt_sd["state"] else None
    if any_step is not None:
        for j, shape in zip((159, 160), FC_SHAPES.values()):
            state[j] = {"step": any_step.clone() if torch.is_tensor(any_step) else any_step,
                        "exp_avg": torch.zeros(shape), "exp_avg_sq": torch.zeros(shape)}
    groups = []
    for gi, g in enumerate(opt_sd["param_groups"]):
        g = dict(g)
        if gi == 0:
            g["params"] = list(range(N_BACKBONE_PARAMS_REF))
        else:
            g["params"] = [p + 2 for p in g["params"]]
        groups.append(g)
    return {"state": state, "param_groups": groups}


def optimizer_from_reference(ref_sd: Mapping) -> Dict:
    n0 = len(ref_sd["param_groups"][0]["params"])
    if n0 == N_BACKBONE_PARAMS:          # already in our layout
        return {"state": dict(ref_sd["state"]), "param_groups": [dict(g) for g in ref_sd["param_groups"]]}
    state = {}
    for idx, st in ref_sd["state"].items():
        idx = int(idx)
        if idx in (159, 160):
            continue
        state[idx if idx < 159 else idx - 2] = st
    groups = []
    for gi, g in enumerate(ref_sd["param_groups"]):
        g = dict(g)
        g["params"] = list(range(N_BACKBONE_PARAMS)) if gi == 0 else [p - 2 for p in g["params"]]
        groups.append(g)
    return {"state": state, "param_groups": groups}


def save_checkpoint(path: str, backbone, decoder, optimizer=None, meta: Optional[Mapping] = None,
                    module_prefix: bool = True) -> None:
    payload = {"backbone": backbone_to_reference(backbone, module_prefix),
               "decoder": decoder_to_reference(decoder, module_prefix)}
    if optimizer is not None:
        payload["optimizer"] = optimizer_to_reference(optimizer.state_dict())
    if meta is not None:
        payload["meta"] = dict(meta)
    tmp = path + ".tmp"
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    torch.save(payload, tmp)
    os.replace(tmp, path)                 # atomic: a crash never leaves a torn checkpoint_latest.pth


def restore_model(model_path, backbone, decoder, optimizer=None, logger=None) -> Dict:
    """Reference ``utils.restore_model`` signature; returns the ``meta`` dict (may be empty)."""
    if model_path is None:
        if logger:
            logger.info("Not using pre-trained model...")
        return {}
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"Model {model_path} does not exist!")
    ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
    for key, loader, model in (("backbone", load_backbone, backbone), ("decoder", load_decoder, decoder)):
        if key in ckpt and model is not None:
            res = loader(model, ckpt[key])
            if logger:
                logger.info("[MODEL_RESTORE] missing keys in %s checkpoint: %s" % (key, set(res.missing_keys)))
                logger.info("[MODEL_RESTORE] missing keys in %s model: %s" % (key, set(res.unexpected_keys)))
    if optimizer is not None and "optimizer" in ckpt:
        optimizer.load_state_dict(optimizer_from_reference(ckpt["optimizer"]))
    meta = dict(ckpt.get("meta", {}))
    if meta:
        # only a checkpoint that carries optimizer state is a resumable training state; weights-only files
        # (``checkpoint_%012d.pth``, upstream releases) are fine-tuning starting points
        meta["has_optimizer"] = "optimizer" in ckpt
    return meta
