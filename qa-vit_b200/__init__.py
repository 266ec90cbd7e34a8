"""qavit_b200: B200-native (sm_100a) hot path of QA-ViT / HQA-ViT behind the reference's nn.Module interface.

The directory is named ``qa-vit_b200`` (not importable as such); ``import qavit_b200`` resolves here through the
shim package at the repo root."""
from ._lib import EXPORTS, LIB_PATH, lib  # noqa: F401  (raises ImportError when the CUDA extension is not built)
from . import functional  # noqa: F401
from .functional import batch_mix, cross_entropy, dropout, manual_seed, mix_batch  # noqa: F401
from .modules import (HQAViT, HQAViTConfig, PatchEmbed, QAViT, QAViTConfig, QuadAttentionBlock,  # noqa: F401
                      QuadBlockWithTokenLearner)
from .optim import FusedAdamW, GradientMonitor, ModelEMA, clip_grad_norms_  # noqa: F401
from .dp import GradAllReducer  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from .transfer import adjust_positional_embedding, load_pretrained_except_head  # noqa: F401
from .evalutil import normalize_batch, tta_views, validate_tta  # noqa: F401

def HQAViTv2(config, **kw):
    """HQAViTv2_CIFAR100.py's ``HQAViT`` (ConvNeXt-patchify CNN stem with LayerScale, V:753-833) -- ``HQAViT(config, variant="v2")``."""
    return HQAViT(config, variant="v2", **kw)


__all__ = ["QAViT", "HQAViT", "HQAViTv2", "QAViTConfig", "HQAViTConfig", "QuadAttentionBlock", "QuadBlockWithTokenLearner",
           "PatchEmbed", "cross_entropy", "FusedAdamW", "clip_grad_norms_", "GradAllReducer", "GraphedTrainStep"]
