"""On-disk formats of the reference's training loop, so that its checkpoints and learning-curve files move
between the two implementations unchanged.

* checkpoint: `torch.save({'model': model.state_dict(), 'optimizer': optimizer.state_dict()}, path)`
  (utils/early_stopping.py:53-65 for the best-validation checkpoint, utils/mnist_utils/train.py:243-247 for
  the end-of-training one).  The vael_b200 models keep the circuit arrays and logic tables out of the state
  dict (non-persistent buffers / plain attributes, as the reference's `w_q`, `herbrand_base`, `W_*` are plain
  attributes: models/vael.py:25,157,279-283), so the keys are exactly `encoder.* / decoder.* / mlp.*`.
* learning curves: `np.save('train_info.npy', {key: [per-epoch mean, ...]})` with the keys of the per-epoch
  dicts (utils/mnist_utils/train.py:74-93,160-164,239-240): elbo, recon_loss, kl_div, queryBCE, labelBCE.
"""
from __future__ import annotations

import os
from typing import Dict, List, Mapping, Optional

import numpy as np
import torch

CURVE_KEYS = ("elbo", "recon_loss", "kl_div", "queryBCE", "labelBCE")   # utils/mnist_utils/train.py:74-93


def save_checkpoint(model: torch.nn.Module, optimizer: torch.optim.Optimizer, path: str) -> str:
    """utils/early_stopping.py:60-63: the two state dicts under 'model' and 'optimizer'."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict()}, path)
    return path


def load_checkpoint(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
                    map_location=None) -> Mapping:
    """Load a checkpoint written by either implementation.  `strict=True`: a key mismatch is an error,
    not something to paper over."""
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    if not isinstance(ckpt, dict) or "model" not in ckpt:
        raise ValueError(f"{path}: not a {{'model', 'optimizer'}} checkpoint")
    model.load_state_dict(ckpt["model"], strict=True)
    if optimizer is not None:
        if "optimizer" not in ckpt:
            raise ValueError(f"{path}: checkpoint holds no optimizer state")
        optimizer.load_state_dict(ckpt["optimizer"])
    return ckpt


def append_epoch(info: Dict[str, List[float]], epoch_terms: Mapping[str, float]) -> Dict[str, List[float]]:
    """utils/mnist_utils/train.py:160-164: per-epoch means appended key by key."""
    for key, value in epoch_terms.items():
        info.setdefault(key, []).append(float(value))
    return info


def save_curves(folder: str, train_info: Mapping[str, List[float]], validation_info: Mapping[str, List[float]]) -> None:
    """utils/mnist_utils/train.py:239-240: the two dicts as pickled 0-d object arrays."""
    os.makedirs(folder, exist_ok=True)
    np.save(os.path.join(folder, "train_info.npy"), dict(train_info))
    np.save(os.path.join(folder, "validation_info.npy"), dict(validation_info))


def load_curves(folder: str):
    """What the reference's analysis code does with these files: `np.load(..., allow_pickle=True).item()`."""
    return tuple(np.load(os.path.join(folder, n), allow_pickle=True).item()
                 for n in ("train_info.npy", "validation_info.npy"))
