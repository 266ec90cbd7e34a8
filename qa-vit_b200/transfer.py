"""Checkpoint / transfer helpers the reference's fine-tune scripts apply to a model.  Host-side, one-off
tensor surgery on the module tree -- not on the hot path -- kept signature-compatible so those scripts run on the
drop-in modules unchanged."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def adjust_positional_embedding(model: nn.Module, new_img_size: int) -> bool:
    """``adjust_positional_embedding`` of HQAViT_Tiny_stl10.py:250-282: bicubic resize of ``model.pos_embed`` [1, N, d] to
    the patch grid of ``new_img_size`` (repeat / truncate when a grid is not square).  Returns True when it changed."""
    # the reference probes patch_embed.patch_size and falls back to 4 (its PatchEmbed has no such attribute, H:1129-1134)
    ps = getattr(model.patch_embed, "patch_size", 4)
    patch = ps[0] if isinstance(ps, (tuple, list)) else int(ps)
    new_n = (new_img_size // patch) ** 2
    pos = model.pos_embed
    old_n = pos.shape[1]
    if new_n == old_n:
        return False
    old_s, new_s = int(math.sqrt(old_n)), int(math.sqrt(new_n))
    if old_s * old_s == old_n and new_s * new_s == new_n:
        grid = pos.detach().reshape(1, old_s, old_s, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=(new_s, new_s), mode="bicubic", align_corners=False)
        new = grid.permute(0, 2, 3, 1).reshape(1, new_n, -1)
    elif new_n > old_n:
        new = pos.detach().repeat(1, new_n // old_n + 1, 1)[:, :new_n, :]
    else:
        new = pos.detach()[:, :new_n, :]
    model.pos_embed = nn.Parameter(new.contiguous())
    return True


def load_pretrained_except_head(model: nn.Module, state_dict: Dict[str, torch.Tensor]) -> Tuple[int, int]:
    """The head-swap load of HQAViT_Tiny_Cifar10.py:445-453 / HQAViT_C100_Finetune.py: every checkpoint tensor whose key
    exists in ``model`` and does not contain 'head' is loaded, the (re-sized) head keeps its fresh initialisation.
    Returns (tensors loaded, tensors skipped)."""
    own = model.state_dict()
    picked = {k: v for k, v in state_dict.items() if k in own and "head" not in k}
    own.update(picked)
    model.load_state_dict(own)
    return len(picked), len(state_dict) - len(picked)
