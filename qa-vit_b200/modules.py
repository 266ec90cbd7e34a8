"""Drop-in nn.Module trees for the reference's QAViT / QAViTv2 / HQAViT classes.

Same constructor arguments (a duck-typed config dataclass), same attribute / parameter names -- hence the same
``state_dict()`` keys, including the GlobalTokenBank entries aliased under every branch -- and the same
``forward(x[B, C, S, S]) -> logits[B, classes]``.  The encoder blocks, patch embedding and
head run through the hand-written sm_100a kernels behind include/qavit_b200.h; there is no eager fallback.
HQAViT's CNN lateral path (cnn_stem / lmfa / rrcv) is one native call per direction (qavit_lateral_*), SplitFusion
another (qavit_splitfusion_*); their sub-modules only hold parameters under the reference's names.

Reference: HQAViT_CIFAR100.py (H), QAViT.py, QAViTv2.py, QAViTv2_CIFAR100.py, HQAViT_IN_Tiny.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as QF
from ._lib import PARAMS, QP, SPLITFUSION_PARAMS, BlockCfg, LateralCfg, SplitFusionCfg, lateral_param_names


# --------------------------------------------------------------------------------------------- configs
@dataclass
class QAViTConfig:
    """Field-for-field the reference's QAViTConfig (QAViTv2_CIFAR100.py:41-60 defaults)."""
    img_size: int = 32
    patch_size: int = 4
    in_channels: int = 3
    num_classes: int = 100
    embed_dim: int = 192
    depth: int = 8
    num_heads: int = 4
    compress_ratio: int = 4
    bottleneck_ratio: int = 2
    mlp_ratio: float = 0.5
    global_bank_size: int = 16
    dropout: float = 0.1
    drop_path: float = 0.1
    window_size: int = 4
    dilation_factors: Tuple[int, ...] = (1, 2)
    landmark_pooling_stride: int = 2
    num_channel_groups: int = 6
    linformer_k: int = 32


@dataclass
class HQAViTConfig(QAViTConfig):
    """Field-for-field the reference's HQAViTConfig (H:42-78)."""
    cnn_c2: int = 64
    cnn_c3: int = 128
    cnn_c4: int = 256
    rrcv_channels: int = 64
    rrcv_num_blocks: int = 1
    use_token_learner: bool = True
    num_learned_tokens: int = 16
    fusion_stages: Tuple[int, ...] = (2, 3, 4)


def _get(cfg, name, default):
    return getattr(cfg, name, default)


# --------------------------------------------------------------------------------------------- parameter holders
class _Holder(nn.Module):
    """Sub-modules of a quad block only hold parameters under the reference's names; the arithmetic of the whole
    block is one native call (QuadAttentionBlock.forward)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is fused into QuadAttentionBlock's native forward; call the block")


class GlobalTokenBank(_Holder):
    """H:275-321 (v1: QAViT.py:183-224 -- no update counter)."""

    def __init__(self, bank_size: int, embed_dim: int, v1: bool = False):
        super().__init__()
        self.bank_size, self.embed_dim, self.v1 = bank_size, embed_dim, v1
        self.global_k = nn.Parameter(torch.randn(1, bank_size, embed_dim) * 0.02)
        self.global_v = nn.Parameter(torch.randn(1, bank_size, embed_dim) * 0.02)
        self.write_norm = nn.LayerNorm(embed_dim)
        self.write_compression = nn.Linear(embed_dim, embed_dim)
        self.write_gate = nn.Linear(embed_dim, bank_size)
        if not v1:
            self.register_buffer("update_count", torch.tensor(0))

    def read(self, batch_size: int):
        return self.global_k.expand(batch_size, -1, -1), self.global_v.expand(batch_size, -1, -1)


class LinformerCompression(_Holder):
    def __init__(self, seq_len: int, compressed_len: int):
        super().__init__()
        self.seq_len, self.compressed_len = seq_len, compressed_len
        self.E_k = nn.Parameter(torch.randn(seq_len, compressed_len) * 0.02)
        self.E_v = nn.Parameter(torch.randn(seq_len, compressed_len) * 0.02)


class EfficientSpatialWindowAttention(_Holder):
    def __init__(self, config, global_bank):
        super().__init__()
        self.global_bank = global_bank
        d = config.embed_dim
        self.qkv = nn.Linear(d, 3 * d, bias=True)
        self.linformer = LinformerCompression(config.window_size ** 2, config.linformer_k)
        self.proj = nn.Linear(d, d)
        self.dropout = nn.Dropout(config.dropout)
        self.norm = nn.LayerNorm(d)


class EfficientMultiScaleDilatedAttention(_Holder):
    def __init__(self, config, global_bank):
        super().__init__()
        self.global_bank = global_bank
        d = config.embed_dim
        self.qkv = nn.Linear(d, 3 * d, bias=True)
        self.linformer = LinformerCompression(128, config.linformer_k)
        self.proj = nn.Linear(d, d)
        self.dropout = nn.Dropout(config.dropout)
        self.norm = nn.LayerNorm(d)


class EfficientChannelGroupAttention(_Holder):
    def __init__(self, config, global_bank):
        super().__init__()
        self.global_bank = global_bank
        d, G = config.embed_dim, config.num_channel_groups
        cg, cpg = d // G, (d // 2) // G
        self.q_proj = nn.Linear(cg, cpg)
        self.k_proj = nn.Linear(cg, cpg)
        self.v_proj = nn.Linear(cg, cpg)
        self.bank_k_proj = nn.Linear(d, cpg)
        self.bank_v_proj = nn.Linear(d, cpg)
        self.proj = nn.Linear(d // 2, d)
        self.dropout = nn.Dropout(config.dropout)
        self.norm = nn.LayerNorm(d)


class CrossAttentionBranch(_Holder):
    def __init__(self, config, global_bank):
        super().__init__()
        self.global_bank = global_bank
        d = config.embed_dim
        self.q_proj = nn.Linear(d, d)
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.proj = nn.Linear(d, d)
        self.dropout = nn.Dropout(config.dropout)


class HybridFusion(_Holder):
    def __init__(self, embed_dim, num_branches=4):
        super().__init__()
        self.fusion_weights = nn.Parameter(torch.ones(num_branches))


class BottleneckMLP(_Holder):
    def __init__(self, input_dim, hidden_dim, output_dim, dropout=0.1):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)


class DepthwiseConv2d(_Holder):
    def __init__(self, dim, kernel_size=3, bias=False, with_scale=True):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size, padding=kernel_size // 2, groups=dim, bias=bias)
        if with_scale:
            self.scale = nn.Parameter(torch.ones(1, dim, 1, 1) * 0.1)


class CCFFFN(_Holder):
    """variant 'v1' = QAViT.py:553-582, 'v2' = H:678-712 (dwconv bias only in QAViTv2.py:856-862)."""

    def __init__(self, embed_dim, mlp_ratio=0.5, dropout=0.1, variant="v2", dwconv_bias=False):
        super().__init__()
        hidden = int(embed_dim * mlp_ratio)
        self.fc1 = nn.Linear(embed_dim, hidden)
        if variant == "v2":
            self.dwconv_norm = nn.LayerNorm(hidden)
        self.dwconv = DepthwiseConv2d(hidden, 3, bias=dwconv_bias, with_scale=variant == "v2")
        if variant == "v2":
            self.post_dwconv_norm = nn.LayerNorm(hidden)
        self.fc2 = nn.Linear(hidden, embed_dim)
        if variant == "v2":
            self.gamma = nn.Parameter(torch.ones(1) * 0.1)


_NO_GRAD_SUFFIX = ("swa.norm.", "msda.norm.", "cga.norm.", "write_norm.", "write_compression.", "write_gate.")


class QuadAttentionBlock(nn.Module):
    """H:1037-1085.  ``variant``: 'v2' (HQAViT, QAViTv2_CIFAR100, QAViTV2_EXTREME), 'v2b' (QAViTv2.py: + dwconv bias),
    'v1' (QAViT.py)."""

    def __init__(self, config, global_bank, drop_path=0., variant: str = "v2"):
        super().__init__()
        self.config = config
        self.variant = variant
        self.drop_path_rate = float(drop_path)
        d = config.embed_dim
        self.embed_dim = d
        self.compressed_dim = d // config.compress_ratio
        self.norm1 = nn.LayerNorm(d)
        self.swa = EfficientSpatialWindowAttention(config, global_bank)
        self.msda = EfficientMultiScaleDilatedAttention(config, global_bank)
        self.cga = EfficientChannelGroupAttention(config, global_bank)
        self.cross_attn = CrossAttentionBranch(config, global_bank)
        self.norm_swa = nn.LayerNorm(d)
        self.norm_msda = nn.LayerNorm(d)
        self.norm_cga = nn.LayerNorm(d)
        self.norm_cross = nn.LayerNorm(d)
        self.compress_swa = nn.Linear(d, self.compressed_dim)
        self.compress_msda = nn.Linear(d, self.compressed_dim)
        self.compress_cga = nn.Linear(d, self.compressed_dim)
        self.compress_cross = nn.Linear(d, self.compressed_dim)
        self.fusion = HybridFusion(self.compressed_dim, 4)
        self.bottleneck_mlp = BottleneckMLP(4 * self.compressed_dim, d // config.bottleneck_ratio, d, config.dropout)
        self.norm2 = nn.LayerNorm(d)
        self.ccf_ffn = CCFFFN(d, config.mlp_ratio, config.dropout, variant="v1" if variant == "v1" else "v2",
                              dwconv_bias=variant in ("v1", "v2b"))
        self.precision = "auto"
        object.__setattr__(self, "_bank_ref", global_bank)   # not a child: registered under the branches already

    def forward(self, x):
        return _block_apply(self, None, x)


class TokenLearner(_Holder):
    def __init__(self, in_dim: int, num_out_tokens: int = 16):
        super().__init__()
        self.num_out_tokens = num_out_tokens
        self.attention = nn.Sequential(nn.LayerNorm(in_dim), nn.Linear(in_dim, num_out_tokens))


class TokenUpMix(_Holder):
    def __init__(self, embed_dim: int, num_in_tokens: int, num_out_tokens: int):
        super().__init__()
        self.upsample_attn = nn.Linear(num_in_tokens, num_out_tokens)
        self.norm = nn.LayerNorm(embed_dim)


class QuadBlockWithTokenLearner(nn.Module):
    """H:1091-1123 (the TinyIN variant rounds num_learned_tokens to a square, HQAViT_IN_Tiny.py:1322-1330)."""

    def __init__(self, config, global_bank, drop_path=0., use_token_learner=True, square_tokens=False):
        super().__init__()
        self.use_token_learner = use_token_learner
        if use_token_learner:
            M = config.num_learned_tokens
            if square_tokens:
                sq = int(math.sqrt(M))
                if sq * sq != M:
                    M = max(4, sq * sq)
            self.token_learner = TokenLearner(config.embed_dim, M)
            self.token_upmix = TokenUpMix(config.embed_dim, M, (config.img_size // config.patch_size) ** 2)
        self.quad_block = QuadAttentionBlock(config, global_bank, drop_path)

    def forward(self, x):
        return _block_apply(self.quad_block, self if self.use_token_learner else None, x)


def _block_apply(block: QuadAttentionBlock, wrapper: Optional[QuadBlockWithTokenLearner], x: torch.Tensor):
    """Assemble the cfg struct + parameter table and run the native block (forward + autograd backward)."""
    cfg = block.config
    bank = block._bank_ref
    B, N, d = x.shape
    c = BlockCfg()
    c.batch, c.tokens_full = B, N
    c.token_learner = 1 if wrapper is not None else 0
    c.tokens = wrapper.token_learner.num_out_tokens if wrapper is not None else N
    # TokenUpMix's Linear fixes the number of tokens it emits at construction (H:1099-1103); it differs from N when a model built
    # for 32 x 32 is fed larger images after adjust_positional_embedding (STL-10 recipe: the first block maps 576 -> 16 -> 64)
    c.tokens_out = wrapper.token_upmix.upsample_attn.out_features if wrapper is not None else 0
    c.dim, c.heads, c.bank_size, c.groups = d, cfg.num_heads, cfg.global_bank_size, cfg.num_channel_groups
    c.window, c.linformer_k, c.msda_seq_len = cfg.window_size, cfg.linformer_k, block.msda.linformer.seq_len
    dil = tuple(cfg.dilation_factors)
    c.n_dilations = len(dil)
    for i, v in enumerate(dil[:4]):
        c.dilations[i] = int(v)
    c.pool_stride = cfg.landmark_pooling_stride
    c.compress_dim = block.compressed_dim
    c.bottleneck_hidden = block.bottleneck_mlp.fc1.out_features
    c.ffn_hidden = block.ccf_ffn.fc1.out_features
    c.ffn_v1 = 1 if block.variant == "v1" else 0
    c.dwconv_bias = 1 if block.variant in ("v1", "v2b") else 0
    c.bank_v1 = 1 if bank.v1 else 0
    c.train = 1 if block.training else 0
    c.dtype = QF.resolve_dtype(block.precision)
    # every nn.Dropout of the block is built from config.dropout (H:416, 486, 556, 611, 1064, 1068); read the live module so
    # that `.p` edits behave like the reference's
    c.dropout = float(block.swa.dropout.p)
    c.drop_path = float(block.drop_path_rate)

    tensors, index, no_grad = [], [], []
    for qi, (name, scope) in enumerate(PARAMS):
        owner = block if scope == 0 else (wrapper if scope == 1 else bank)
        if owner is None:
            continue
        obj = owner
        try:
            for part in name.split("."):
                obj = getattr(obj, part)
        except AttributeError:
            continue                     # parameter not present in this variant (e.g. gamma in v1)
        if obj is None:
            continue
        tensors.append(obj)
        index.append(qi)
        if any(sfx in name + "." or name.startswith(sfx) for sfx in _NO_GRAD_SUFFIX):
            no_grad.append(qi)
    meta = QF.BlockMeta(c, index, no_grad, None if bank.v1 else bank.update_count)
    return QF.QuadBlockFn.apply(x, meta, *tensors)


# --------------------------------------------------------------------------------------------- patch embed / models
class _PatchConv(nn.Conv2d):
    """``patch_embed.proj``: an nn.Conv2d (same parameters / state_dict keys) whose own forward -- used only when someone calls
    the conv itself or hangs hooks on it (Grad-CAM registers a forward hook here, test_hqa.py:241-259) -- runs the library's
    GEMM on the gathered patches (stride == kernel, so the gather is a reshape) and returns [B, d, H/p, W/p]."""

    def forward(self, x):
        B, Cin, S, S2 = x.shape
        p = self.kernel_size[0]
        if self.stride[0] != p or S % p or S2 % p:
            raise RuntimeError("qavit_b200: patch projection needs stride == kernel and an image divisible by the patch")
        hp, wp = S // p, S2 // p
        patches = x.reshape(B, Cin, hp, p, wp, p).permute(0, 2, 4, 1, 3, 5).reshape(B * hp * wp, Cin * p * p)
        y = QF.linear(patches, self.weight, self.bias)                       # [B hp wp, d]
        return y.view(B, hp, wp, -1).permute(0, 3, 1, 2)                      # channels-last view of [B, d, hp, wp]


class PatchEmbed(nn.Module):
    """H:1129-1138.  One native call (gather + GEMM + LayerNorm + pos) -- unless ``proj`` carries hooks: then the projection
    runs as its own autograd node through ``proj.__call__`` so that forward / backward hooks fire with the reference's
    [B, d, H/p, W/p] activation (Grad-CAM, test_hqa.py:259), followed by the native LayerNorm."""

    def __init__(self, img_size=32, patch_size=4, in_channels=3, embed_dim=192):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = _PatchConv(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)
        self.precision = "auto"

    def _hooked(self) -> bool:
        pj = self.proj
        return bool(pj._forward_hooks or pj._forward_pre_hooks or pj._backward_hooks or getattr(pj, "_backward_pre_hooks", None))

    def forward(self, x, pos: Optional[torch.Tensor] = None):
        if self._hooked():
            y = self.proj(x)                                                 # hooks fire here
            t = y.flatten(2).transpose(1, 2)                                 # H:1137
            t = QF.LayerNormFn.apply(t, self.norm.weight, self.norm.bias, self.norm.eps)
            return t if pos is None else t + pos
        return QF.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias, pos,
                                     QF.resolve_dtype(self.precision))


def _init_weights(m):
    """The reference's _init_weights (H:1215-1224)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _Base(nn.Module):
    precision = "auto"
    # The reference's head is an nn.Linear, so under torch.autocast(bfloat16) its logits come back as bf16.  The
    # native head computes them in fp32; set this to True to round them to the autocast dtype for bit-for-bit dtype parity
    # with scripts that inspect ``outputs.dtype`` -- off by default because it only loses precision.
    match_autocast_output_dtype = False

    def _finish_logits(self, logits):
        if self.match_autocast_output_dtype and torch.is_autocast_enabled():
            return logits.to(torch.get_autocast_gpu_dtype())
        return logits

    def set_precision(self, precision: str):
        """'auto' (bf16 under torch.autocast, else fp32), 'fp32' or 'bf16' for every native block."""
        assert precision in ("auto", "fp32", "bf16")
        self.precision = precision
        for m in self.modules():
            if isinstance(m, (QuadAttentionBlock, SplitFusion, PatchEmbed)):
                m.precision = precision
        return self

    def _stream_dropout(self, T):
        p = float(self.pos_drop.p)
        return QF.dropout(T, p, True) if (self.training and p > 0) else T


class QAViT(_Base):
    """QAViT.py:654-699 / QAViTv2.py:1011-1055.  ``variant``: 'v1' = QAViT.py, 'v2b' = QAViTv2.py (dwconv bias),
    'v2' = QAViTv2_CIFAR100.py / QAViTV2_EXTREME.py."""

    def __init__(self, config, variant: str = "v2"):
        super().__init__()
        self.config = config
        self.variant = variant
        d = config.embed_dim
        self.num_patches = (config.img_size // config.patch_size) ** 2
        self.patch_embed = PatchEmbed(config.img_size, config.patch_size, config.in_channels, d)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, d))
        self.pos_drop = nn.Dropout(config.dropout)
        self.global_bank = GlobalTokenBank(config.global_bank_size, d, v1=variant == "v1")
        dpr = [v.item() for v in torch.linspace(0, config.drop_path, config.depth)]
        self.blocks = nn.ModuleList([QuadAttentionBlock(config, self.global_bank, dpr[i], variant) for i in range(config.depth)])
        self.norm = nn.LayerNorm(d)
        self.head = nn.Linear(d, config.num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        self.apply(_init_weights)

    def forward(self, x):
        T = self.patch_embed(x, self.pos_embed)
        T = self._stream_dropout(T)
        for blk in self.blocks:
            T = blk(T)
        return self._finish_logits(QF.HeadFn.apply(T, self.norm.weight, self.norm.bias, self.head.weight, self.head.bias))


# --------------------------------------------------------------------------------------------- HQAViT lateral path (row f-1)
class _LateralHolder(nn.Module):
    """The lateral modules hold parameters / buffers under the reference's names; their arithmetic is part of ONE native
    call per direction (HQAViT.lateral -> qavit_lateral_forward / backward)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is fused into HQAViT.lateral()'s native call (qavit_lateral_forward)")


class ConvNeXtBlock(_LateralHolder):
    """H:718-739."""

    def __init__(self, dim, drop_path=0.):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)


class ConvNeXtBlockV2(_LateralHolder):
    """HQAViTv2_CIFAR100.py:718-751: ConvNeXt block with LayerScale ``gamma`` and DropPath (``drop_path_rate``)."""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim))
        self.drop_path_rate = float(drop_path)


class CNNStemModelV2(_LateralHolder):
    """HQAViTv2_CIFAR100.py:753-833: 4x4 patchify stem + LayerNorm([c2, g, g]); stages of 2 / 3 / 2 LayerScale ConvNeXt blocks at
    c2 / c3 / c4 channels on the g x g map, LayerNorm([C, g, g]) + 1x1 conv between them.  (The reference hard-codes g = 8.)"""

    def __init__(self, in_ch=3, c2=64, c3=128, c4=256, grid=8):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_ch, c2, kernel_size=4, stride=4), nn.LayerNorm([c2, grid, grid], eps=1e-6))
        self.stage2 = nn.Sequential(ConvNeXtBlockV2(c2, 0.0), ConvNeXtBlockV2(c2, 0.0))
        self.downsample2 = nn.Sequential(nn.LayerNorm([c2, grid, grid], eps=1e-6), nn.Conv2d(c2, c3, kernel_size=1))
        self.stage3 = nn.Sequential(ConvNeXtBlockV2(c3, 0.0), ConvNeXtBlockV2(c3, 0.1), ConvNeXtBlockV2(c3, 0.1))
        self.downsample3 = nn.Sequential(nn.LayerNorm([c3, grid, grid], eps=1e-6), nn.Conv2d(c3, c4, kernel_size=1))
        self.stage4 = nn.Sequential(ConvNeXtBlockV2(c4, 0.1), ConvNeXtBlockV2(c4, 0.1))

    def blocks_in_order(self):
        return [*self.stage2, *self.stage3, *self.stage4]


class CNNStemModel(_LateralHolder):
    """H:742-793."""

    def __init__(self, in_ch=3, c2=64, c3=128, c4=256):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_ch, 32, 3, stride=2, padding=1), nn.BatchNorm2d(32), nn.GELU())
        self.stage1 = nn.Sequential(nn.Conv2d(32, c2, 3, stride=2, padding=1), nn.BatchNorm2d(c2), nn.GELU(), ConvNeXtBlock(c2))
        self.stage2 = nn.Sequential(nn.Conv2d(c2, c3, 1), nn.BatchNorm2d(c3), ConvNeXtBlock(c3))
        self.stage3 = nn.Sequential(nn.Conv2d(c3, c4, 1), nn.BatchNorm2d(c4), ConvNeXtBlock(c4))


class LMFAdapter(_LateralHolder):
    """H:799-849."""

    def __init__(self, in_channels: int, embed_dim: int, target_hw: int = 8):
        super().__init__()
        self.target_hw = target_hw
        self.dwconv_3x3 = nn.Conv2d(in_channels, in_channels, 3, padding=1, groups=in_channels)
        self.dwconv_5x5 = nn.Conv2d(in_channels, in_channels, 5, padding=2, groups=in_channels)
        self.proj = nn.Conv2d(3 * in_channels, embed_dim, 1)
        self.norm = nn.LayerNorm(embed_dim)
        self.act = nn.GELU()


class RRCV(_LateralHolder):
    """H:855-907."""

    def __init__(self, embed_dim: int, rec_channels: int = 64, num_blocks: int = 1, layer_scale: bool = False):
        super().__init__()
        self.reverse_proj = nn.Conv2d(embed_dim, rec_channels, 1)
        blk = ConvNeXtBlockV2 if layer_scale else ConvNeXtBlock       # HQAViTv2's RRCV blocks carry LayerScale (V:718-751, 895)
        self.blocks = nn.ModuleList([blk(rec_channels) for _ in range(num_blocks)])
        self.reembed_proj = nn.Conv2d(rec_channels, embed_dim, 1)
        self.norm = nn.LayerNorm(embed_dim)
        self.beta = nn.Parameter(torch.tensor(0.1))


class SplitFusion(nn.Module):
    """H:913-965 (keeps the hard-coded Dropout(0.1) of H:930 as cat_mlp[3]); one native call per direction."""

    def __init__(self, embed_dim: int):
        super().__init__()
        self.gate_norm = nn.LayerNorm(embed_dim)
        self.gate_fc = nn.Linear(embed_dim, embed_dim)
        self.cat_mlp = nn.Sequential(nn.Linear(2 * embed_dim, embed_dim), nn.LayerNorm(embed_dim), nn.GELU(), nn.Dropout(0.1))
        self.fusion_weights = nn.Parameter(torch.tensor([0.75, 0.25]))
        self.final_norm = nn.LayerNorm(embed_dim)
        self.precision = "auto"

    def forward(self, T_in, R):
        B, N, d = T_in.shape
        c = SplitFusionCfg()
        c.rows, c.dim = B * N, d
        c.dtype = QF.resolve_dtype(self.precision)
        c.train = 1 if self.training else 0
        c.drop_p = float(self.cat_mlp[3].p)
        tensors = [_resolve(self, n) for n in SPLITFUSION_PARAMS]
        return QF.SplitFusionFn.apply(T_in, R, c, *tensors)


def _resolve(root, dotted):
    obj = root
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


class HQAViT(_Base):
    """H:1141-1277.  ``stage_depths`` = (2, 2, 2, 2) for CIFAR-100, (2, 2, 6, 2) for TinyImageNet
    (HQAViT_IN_Tiny.py:1399-1420; that file also forces a square number of learned tokens)."""

    def __init__(self, config, stage_depths: Optional[Tuple[int, ...]] = None, square_tokens: bool = False, variant: str = "v1"):
        super().__init__()
        assert variant in ("v1", "v2"), variant      # v2: HQAViTv2_CIFAR100.py (same model around a ConvNeXt-patchify stem)
        self.variant = variant
        self.config = config
        d = config.embed_dim
        if stage_depths is None:
            stage_depths = (2, 2, 2, 2) if config.depth == 8 else (2, 2, config.depth - 6, 2)
        assert sum(stage_depths) == config.depth and len(stage_depths) == 4
        self.stage_depths = tuple(stage_depths)
        self.num_patches = (config.img_size // config.patch_size) ** 2
        self.H = self.W = config.img_size // config.patch_size
        self.patch_embed = PatchEmbed(config.img_size, config.patch_size, config.in_channels, d)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, d))
        self.pos_drop = nn.Dropout(config.dropout)
        self.global_bank = GlobalTokenBank(config.global_bank_size, d)
        if variant == "v2":
            self.cnn_stem = CNNStemModelV2(config.in_channels, config.cnn_c2, config.cnn_c3, config.cnn_c4, grid=config.img_size // 4)
        else:
            self.cnn_stem = CNNStemModel(config.in_channels, config.cnn_c2, config.cnn_c3, config.cnn_c4)
        self.lmfa2 = LMFAdapter(config.cnn_c2, d, target_hw=self.H)
        self.lmfa3 = LMFAdapter(config.cnn_c3, d, target_hw=self.H)
        self.lmfa4 = LMFAdapter(config.cnn_c4, d, target_hw=self.H)
        self.rrcv2 = RRCV(d, config.rrcv_channels, config.rrcv_num_blocks, variant == "v2")
        self.rrcv3 = RRCV(d, config.rrcv_channels, config.rrcv_num_blocks, variant == "v2")
        self.rrcv4 = RRCV(d, config.rrcv_channels, config.rrcv_num_blocks, variant == "v2")
        self.fuse2 = SplitFusion(d)
        self.fuse3 = SplitFusion(d)
        self.fuse4 = SplitFusion(d)
        dpr = [v.item() for v in torch.linspace(0, config.drop_path, config.depth)]
        k = 0
        for st, n in enumerate(self.stage_depths, start=1):
            blocks = nn.ModuleList([
                QuadBlockWithTokenLearner(config, self.global_bank, dpr[k + i], config.use_token_learner, square_tokens)
                for i in range(n)])
            setattr(self, f"stage{st}_blocks", blocks)
            k += n
        self.norm = nn.LayerNorm(d)
        self.head = nn.Linear(d, config.num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        self.apply(_init_weights)
        self._lateral_names = None

    def lateral(self, x, phased: bool = False):
        """cnn_stem -> lmfa{2,3,4} -> rrcv{2,3,4} (H:1236-1247) as one native call: image -> (R2, R3, R4).
        phased=True: only the stem runs here -> (state, token); lateral_adapter(state, token, k) then yields R_k."""
        cfg = self.config
        c = LateralCfg()
        c.batch, c.img_size, c.in_channels = x.shape[0], x.shape[-1], x.shape[1]
        v2 = self.variant == "v2"
        rng = None
        if v2:
            c.stem_kind = 1
            c.c_stem = 32
            c.c2, c.c3, c.c4 = (self.cnn_stem.stem[0].out_channels, self.cnn_stem.downsample2[1].out_channels,
                                self.cnn_stem.downsample3[1].out_channels)
            for j, blk in enumerate(self.cnn_stem.blocks_in_order()):
                c.stem_drop_path[j] = float(blk.drop_path_rate)
            if self.training and any(blk.drop_path_rate > 0 for blk in self.cnn_stem.blocks_in_order()):
                # its own Philox state: the lateral path runs on a side stream next to calls that advance the shared one
                if getattr(self, "_stem_rng", None) is None or self._stem_rng.device != x.device:
                    seed = torch.initial_seed() if QF._rng_seed is None else QF._rng_seed
                    object.__setattr__(self, "_stem_rng", torch.tensor([(seed ^ 0x5C5C5C5C) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64,
                                                                       device=x.device))
                rng = self._stem_rng
                c.rng = rng.data_ptr()
        else:
            c.c_stem = self.cnn_stem.stem[0].out_channels
            c.c2, c.c3, c.c4 = (self.cnn_stem.stage1[0].out_channels, self.cnn_stem.stage2[0].out_channels,
                                self.cnn_stem.stage3[0].out_channels)
        c.rrcv_channels, c.rrcv_blocks = self.rrcv2.reverse_proj.out_channels, len(self.rrcv2.blocks)
        c.dim, c.grid = cfg.embed_dim, self.H
        c.train = 1 if self.training else 0
        c.dtype = QF.resolve_dtype(self.precision)
        if not v2:
            bn = self.cnn_stem.stem[1]
            c.bn_eps, c.bn_momentum = float(bn.eps), float(bn.momentum)
        if self._lateral_names is None:
            self._lateral_names = lateral_param_names(c)
        names = self._lateral_names
        tensors = [_resolve(self, n) for n in names]
        buffers = [i for i, n in enumerate(names) if n.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        if not phased:
            return QF.LateralFn.apply(x, QF.LateralMeta(c, buffers), *tensors)
        QF._require_cuda(x, "image batch")
        state = QF.LateralState(c, names, tensors, buffers, x.detach().float().contiguous())
        state.rng_keepalive = rng
        token = QF.LateralPartFn.apply(x, state, 0, *state.part_tensors(0))
        return state, token

    def lateral_adapter(self, state, token, k):
        """LMFAdapter + RRCV of stage k (2..4) on the stem's feature maps -> R_k (phased form of lateral())."""
        return QF.LateralPartFn.apply(token, state, k - 1, *state.part_tensors(k - 1))

    # The lateral path runs on a side stream, in phases: the stem next to patch embed + stage 1, the adapter of stage k issued right
    # before fuse_k -- so adapters 3 / 4 overlap the blocks of stages 2 / 3, and in backward every adapter starts as soon as its dR is
    # known (autograd runs a node's backward on its forward stream) instead of after fuse2's backward.
    concurrent_lateral = os.environ.get("QAVIT_LATERAL_SERIAL", "0") != "1"

    def forward(self, x):
        concurrent = self.concurrent_lateral and x.is_cuda
        if not concurrent:
            R2, R3, R4 = self.lateral(x)
            Rs = {2: R2, 3: R3, 4: R4}
        else:
            if getattr(self, "_lat_stream", None) is None:
                object.__setattr__(self, "_lat_stream", torch.cuda.Stream(device=x.device))
            side = self._lat_stream
            main = torch.cuda.current_stream(x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                state, token = self.lateral(x, phased=True)

        def R_of(k):
            if not concurrent:
                return Rs[k]
            with torch.cuda.stream(side):
                R = self.lateral_adapter(state, token, k)
            R.record_stream(main)
            main.wait_stream(side)
            return R

        T = self.patch_embed(x, self.pos_embed)
        T = self._stream_dropout(T)
        for blk in self.stage1_blocks:
            T = blk(T)
        T = self.fuse2(T, R_of(2))
        for blk in self.stage2_blocks:
            T = blk(T)
        T = self.fuse3(T, R_of(3))
        for blk in self.stage3_blocks:
            T = blk(T)
        T = self.fuse4(T, R_of(4))
        for blk in self.stage4_blocks:
            T = blk(T)
        return self._finish_logits(QF.HeadFn.apply(T, self.norm.weight, self.norm.bias, self.head.weight, self.head.bias))
