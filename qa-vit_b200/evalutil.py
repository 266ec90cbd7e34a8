"""Evaluation-side helpers of the reference's scripts, kept on the device.

``normalize_batch``  transforms.ToTensor() + transforms.Normalize(mean, std) (HQAViT_CIFAR100.py:1300) for a whole batch in one
                     kernel, bit-identical to the torchvision pipeline; ``hflip=True`` is RandomHorizontalFlip(p=1), the second
                     test-time-augmentation view of HQAViT_C100_Finetune.py:110-114.
``validate_tta``     drop-in for HQAViT_C100_Finetune.py:345-384: same arguments, same return value (accuracy in percent of the
                     argmax of the mean softmax over the TTA loaders); probabilities stay on the device and the host is
                     synchronised once per loader instead of once per batch.
``tta_views``        the deterministic views (identity, horizontal flip) generated on the device from ONE uint8 batch.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.nn.functional as F

from ._lib import check, lib

CIFAR100_MEAN = (0.5071, 0.4867, 0.4408)     # HQAViT_CIFAR100.py:1283-1284
CIFAR100_STD = (0.2675, 0.2565, 0.2761)


def normalize_batch(x: torch.Tensor, mean: Sequence[float], std: Sequence[float], hflip: bool = False) -> torch.Tensor:
    """x: uint8 [B, H, W, C] (raw dataset layout), uint8 [B, C, H, W] or float [B, C, H, W] in [0, 1] -> float [B, C, H, W]."""
    if not x.is_cuda:
        raise RuntimeError("qavit_b200.normalize_batch: CUDA tensor expected -- there is no CPU path in this package")
    x = x.contiguous()
    if x.dtype == torch.uint8 and x.dim() == 4 and x.shape[-1] in (1, 3) and x.shape[1] not in (1, 3):
        kind, (B, H, W, C) = 0, x.shape
    elif x.dtype == torch.uint8:
        kind, (B, C, H, W) = 1, x.shape
    elif x.dtype == torch.float32:
        kind, (B, C, H, W) = 2, x.shape
    else:
        raise RuntimeError(f"normalize_batch: dtype {x.dtype} (uint8 or float32)")
    m = torch.tensor(list(mean), dtype=torch.float32, device=x.device)
    s = torch.tensor(list(std), dtype=torch.float32, device=x.device)
    if m.numel() != C or s.numel() != C:
        raise RuntimeError(f"normalize_batch: {C} channels but {m.numel()} means / {s.numel()} stds")
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
    check(lib.qavit_normalize_images(x.data_ptr(), kind, B, C, H, W, m.data_ptr(), s.data_ptr(), int(bool(hflip)), out.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
    return out


def tta_views(x_uint8: torch.Tensor, mean=CIFAR100_MEAN, std=CIFAR100_STD, n: int = 2) -> List[torch.Tensor]:
    """The deterministic test-time-augmentation views of HQAViT_C100_Finetune.py:105-114 from one raw batch on the device:
    [ToTensor + Normalize, HorizontalFlip + ToTensor + Normalize] (the remaining reference views draw random crops / jitter
    in the DataLoader workers and are not reproducible by construction)."""
    views = [normalize_batch(x_uint8, mean, std, hflip=False)]
    if n > 1:
        views.append(normalize_batch(x_uint8, mean, std, hflip=True))
    return views


@torch.no_grad()
def validate_tta(model: torch.nn.Module, tta_loaders: Iterable) -> float:
    """HQAViT_C100_Finetune.py:345-384."""
    model.eval()
    all_predictions = []
    all_targets = None
    for loader_idx, loader in enumerate(tta_loaders):
        batch_predictions, targets_list = [], []
        for inputs, targets in loader:
            outputs = model(inputs.cuda(non_blocking=True))
            batch_predictions.append(F.softmax(outputs.float(), dim=1))        # stays on the device
            if loader_idx == 0:
                targets_list.append(targets)
        if batch_predictions:
            all_predictions.append(torch.cat(batch_predictions, dim=0))
        if loader_idx == 0 and targets_list:
            all_targets = torch.cat([t.to(all_predictions[0].device, non_blocking=True) for t in targets_list], dim=0)
    if len(all_predictions) == 0 or all_targets is None:
        return 0.0
    predicted = torch.stack(all_predictions).mean(dim=0).argmax(dim=1)
    correct = predicted.eq(all_targets).sum().item()
    return 100. * correct / all_targets.size(0)
