"""CPU oracle for the QA-ViT / HQA-ViT hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, in plain fp32 torch tensor arithmetic on
the CPU, of what the reference's nn.Module trees compute.  It is NOT part of the
product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import it.  The product path (qavit_b200) never does.

Parity pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the *live* reference modules imported in the
build container (tests/golden/make_golden.py -> tests/golden/*.npz, committed) and,
when /root/reference is present, against the live modules directly
(tests/test_oracle.py).  See DESIGN.md "Oracle".

Everything is a pure function of (state_dict, cfg, x): parameters are read from a
flat ``{key: tensor}`` dict using the reference's own state_dict key names
(SURVEY.md appendix B), so the same dict loads into the reference, the oracle and
the CUDA modules.  Gradients come from torch autograd over these functions.

Citations are ``file:line`` relative to the reference checkout; H = HQAViT_CIFAR100.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config
@dataclass
class OracleConfig:
    """Union of QAViTConfig (QAViT.py:36-56, QAViTv2.py:43-60) and HQAViTConfig (H:42-78)."""
    family: str = "hqavit"           # hqavit | qavit_v1 | qavit_v2
    img_size: int = 32
    built_img_size: int = 0          # image size the MODULES were built for when it differs from the input (STL-10 fine-tuning feeds
                                     # 96 x 96 images to the 32 x 32 CIFAR model after resizing pos_embed, HQAViT_Tiny_stl10.py:250-282:
                                     # TokenUpMix still emits (32 / 4)^2 tokens and LMFAdapter resizes to 8 x 8, H:839-843); 0 = img_size
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
    window_size: int = 4
    dilation_factors: Tuple[int, ...] = (1, 2)
    landmark_pooling_stride: int = 2
    num_channel_groups: int = 6
    linformer_k: int = 32
    msda_seq_len: int = 128          # hard-coded at H:483
    dwconv_bias: bool = False        # True for QAViT.py / QAViTv2.py (QAViTv2.py:856-862)
    # HQAViT only
    cnn_c2: int = 64
    cnn_c3: int = 128
    cnn_c4: int = 256
    rrcv_channels: int = 64
    rrcv_num_blocks: int = 1
    use_token_learner: bool = True
    num_learned_tokens: int = 16
    stage_depths: Tuple[int, ...] = (2, 2, 2, 2)   # TinyIN: (2, 2, 6, 2)
    stem: str = "v1"                 # hqavit only: "v2" = HQAViTv2_CIFAR100.py's ConvNeXt-patchify stem (V:753-833)


def _ln(x: Tensor, sd, prefix: str, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm: biased variance over the last axis, affine."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _lin(x: Tensor, sd, prefix: str) -> Tensor:
    """nn.Linear: y = x W^T + b, W[out, in]."""
    return x @ sd[prefix + ".weight"].t() + sd[prefix + ".bias"]


def _gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (H:647)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _attn(q: Tensor, k: Tensor, v: Tensor, keep: Optional[Tensor] = None) -> Tensor:
    """efficient_attention (H:355-397) on its SDPA branch: softmax(q k^T / sqrt(hd)) v,
    no mask.  q[..., Nq, hd], k/v[..., Nkv, hd].  ``keep`` (same shape as the
    probabilities, values 0 or 1/(1-p)) is SDPA's dropout_p applied as an explicit
    keep-scale: P <- P * keep, no renormalisation (H:387-392)."""
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(q.shape[-1]))
    P = torch.softmax(s, dim=-1)
    if keep is not None:
        P = P * keep
    return P @ v


def _drop(x: Tensor, masks: Optional[dict], key: str) -> Tensor:
    """nn.Dropout / DropPath with an explicit keep-scale tensor (0 or 1/(1-p)); identity when
    ``masks`` is None or has no entry for ``key``.  The masks are test inputs: torch's own RNG
    stream is not part of the contract, the positions of the dropout sites are."""
    if masks is None or masks.get(key) is None:
        return x
    return x * masks[key]


# --------------------------------------------------------------------------- bank
class Bank:
    """GlobalTokenBank state (H:275-321; v1 QAViT.py:183-224).

    ``k``/``v`` are the leaf parameters (gradients of every read accumulate on
    them); ``off_k``/``off_v`` hold the raw ``.data`` mutation accumulated by
    write(), which autograd never sees (H:315-319)."""

    def __init__(self, sd, cfg: OracleConfig, prefix: str = "global_bank"):
        self.sd, self.cfg, self.p = sd, cfg, prefix
        self.k = sd[prefix + ".global_k"]
        self.v = sd[prefix + ".global_v"]
        self.off_k = torch.zeros_like(self.k)
        self.off_v = torch.zeros_like(self.v)
        self.v1 = cfg.family == "qavit_v1"
        self.count = 0 if self.v1 else int(sd[prefix + ".update_count"])

    def read(self) -> Tuple[Tensor, Tensor]:
        """H:290-294 -- value at this moment, [1, Kb, d]."""
        return self.k + self.off_k, self.v + self.off_v

    @torch.no_grad()
    def write(self, t: Tensor) -> None:
        """H:296-321.  t[B, N, d] is already LN'd by the calling branch's .norm."""
        sd, p = self.sd, self.p
        tn = _ln(t, sd, p + ".write_norm")
        c = _lin(tn, sd, p + ".write_compression")
        g = torch.softmax(_lin(tn, sd, p + ".write_gate"), dim=1)     # over tokens
        uk = (g.transpose(1, 2) @ c).mean(0, keepdim=True)
        uv = (g.transpose(1, 2) @ tn).mean(0, keepdim=True)
        if self.v1:                                                    # QAViT.py:217-224
            uclamp, rate, bclamp = 0.1, 0.01, 1.0
        else:
            uclamp, bclamp = 0.05, 0.5
            rate = 0.005 if self.count < 1000 else 0.01
        uk, uv = uk.clamp(-uclamp, uclamp), uv.clamp(-uclamp, uclamp)
        kd = (self.k + self.off_k + rate * uk).clamp(-bclamp, bclamp)
        vd = (self.v + self.off_v + rate * uv).clamp(-bclamp, bclamp)
        self.off_k = kd - self.k.detach()
        self.off_v = vd - self.v.detach()
        if not self.v1:
            self.count += 1


# --------------------------------------------------------------------------- branches
def _heads(x: Tensor, H: int) -> Tensor:
    """[B, N, H*hd] -> [B, H, N, hd] (head h = channels h*hd .. h*hd+hd-1)."""
    B, N, C = x.shape
    return x.reshape(B, N, H, C // H).transpose(1, 2)


def _linformer(k: Tensor, v: Tensor, E_k: Tensor, E_v: Tensor) -> Tuple[Tensor, Tensor]:
    """LinformerCompression.forward (H:332-352): pad/truncate the token axis to
    E.shape[0] then K' = E_k^T K."""
    L = E_k.shape[0]
    N = k.shape[2]
    if N < L:
        k = F.pad(k, (0, 0, 0, L - N))
        v = F.pad(v, (0, 0, 0, L - N))
    elif N > L:
        k, v = k[:, :, :L], v[:, :, :L]
    return E_k.t() @ k, E_v.t() @ v


def swa(x: Tensor, sd, p: str, cfg: OracleConfig, bank: Bank, train: bool, masks: Optional[dict] = None) -> Tensor:
    """EfficientSpatialWindowAttention.forward (H:441-469)."""
    B, N, C = x.shape
    s = int(math.isqrt(N))
    w, H = cfg.window_size, cfg.num_heads
    assert s * s == N and s % w == 0, "reference requires grid side % window == 0 (H:466)"
    nh = s // w
    xw = x.view(B, nh, w, nh, w, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nh * nh, w * w, C)
    qkv = _lin(xw, sd, p + ".qkv").reshape(-1, w * w, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    kc, vc = _linformer(k, v, sd[p + ".linformer.E_k"], sd[p + ".linformer.E_v"])
    bk, bv = bank.read()
    bk = _heads(bk, H).expand(q.shape[0], -1, -1, -1)
    bv = _heads(bv, H).expand(q.shape[0], -1, -1, -1)
    o = _attn(q, torch.cat([kc, bk], 2), torch.cat([vc, bv], 2), masks.get("att_swa") if masks else None)
    o = o.transpose(1, 2).reshape(-1, w * w, C)
    o = _lin(o, sd, p + ".proj")
    o = o.view(B, nh, nh, w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(B, N, C)
    o = _drop(o, masks, "proj_swa")        # H:465 (elementwise, so applying it after the window reverse is the same op)
    if train:
        bank.write(_ln(o.detach(), sd, p + ".norm"))
    return o


def msda_pool(x: Tensor, cfg: OracleConfig) -> Tensor:
    """Dilated gather (H:489-494), concat, AvgPool1d over the token axis (H:499-501)."""
    B, N, C = x.shape
    s = int(math.isqrt(N))
    g = x.view(B, s, s, C)
    xm = torch.cat([g[:, ::d, ::d, :].reshape(B, -1, C) for d in cfg.dilation_factors], 1)
    st = cfg.landmark_pooling_stride
    n_out = (xm.shape[1] - st) // st + 1
    return xm[:, : n_out * st].reshape(B, n_out, st, C).mean(2)


def msda(x: Tensor, sd, p: str, cfg: OracleConfig, bank: Bank, train: bool, masks: Optional[dict] = None) -> Tensor:
    """EfficientMultiScaleDilatedAttention.forward (H:496-532)."""
    B, N, C = x.shape
    H = cfg.num_heads
    xp = msda_pool(x, cfg)
    kv = _lin(xp, sd, p + ".qkv").reshape(B, -1, 3, H, C // H).permute(2, 0, 3, 1, 4)
    kc, vc = _linformer(kv[1], kv[2], sd[p + ".linformer.E_k"], sd[p + ".linformer.E_v"])
    bk, bv = bank.read()
    bk = _heads(bk, H).expand(B, -1, -1, -1)
    bv = _heads(bv, H).expand(B, -1, -1, -1)
    q = _lin(x, sd, p + ".qkv").reshape(B, N, 3, H, C // H)[:, :, 0].permute(0, 2, 1, 3)
    o = _attn(q, torch.cat([kc, bk], 2), torch.cat([vc, bv], 2), masks.get("att_msda") if masks else None)
    o = _drop(_lin(o.transpose(1, 2).reshape(B, N, C), sd, p + ".proj"), masks, "proj_msda")   # H:529
    if train:
        bank.write(_ln(o.detach(), sd, p + ".norm"))
    return o


def cga(x: Tensor, sd, p: str, cfg: OracleConfig, bank: Bank, train: bool, masks: Optional[dict] = None) -> Tensor:
    """EfficientChannelGroupAttention.forward (H:559-595)."""
    B, N, C = x.shape
    G, H = cfg.num_channel_groups, cfg.num_heads
    cg = C // G
    cpg = (C // 2) // G
    xf = x.view(B, N, G, cg).permute(0, 2, 1, 3).reshape(B * G, N, cg)
    q = _heads(_lin(xf, sd, p + ".q_proj"), H)
    k = _heads(_lin(xf, sd, p + ".k_proj"), H)
    v = _heads(_lin(xf, sd, p + ".v_proj"), H)
    bk, bv = bank.read()
    bk = _heads(_lin(bk, sd, p + ".bank_k_proj"), H).expand(B * G, -1, -1, -1)
    bv = _heads(_lin(bv, sd, p + ".bank_v_proj"), H).expand(B * G, -1, -1, -1)
    o = _attn(q, torch.cat([k, bk], 2), torch.cat([v, bv], 2), masks.get("att_cga") if masks else None)
    o = o.transpose(1, 2).reshape(B, G, N, cpg).permute(0, 2, 1, 3).reshape(B, N, G * cpg)
    o = _drop(_lin(o, sd, p + ".proj"), masks, "proj_cga")                                    # H:592
    if train:
        bank.write(_ln(o.detach(), sd, p + ".norm"))
    return o


def cross(x: Tensor, sd, p: str, cfg: OracleConfig, bank: Bank, masks: Optional[dict] = None) -> Tensor:
    """CrossAttentionBranch.forward (H:613-626)."""
    B, N, C = x.shape
    H = cfg.num_heads
    q = _heads(_lin(x, sd, p + ".q_proj"), H)
    bk, bv = bank.read()
    k = _heads(_lin(bk, sd, p + ".k_proj"), H).expand(B, -1, -1, -1)
    v = _heads(_lin(bv, sd, p + ".v_proj"), H).expand(B, -1, -1, -1)
    o = _attn(q, k, v, masks.get("att_cross") if masks else None).transpose(1, 2).reshape(B, N, C)
    return _drop(_lin(o, sd, p + ".proj"), masks, "proj_cross")                               # H:625


def ccf_ffn(x: Tensor, sd, p: str, cfg: OracleConfig, masks: Optional[dict] = None) -> Tensor:
    """CCFFFN.forward: v2 (H:700-712, QAViTv2.py:864-885) / v1 (QAViT.py:571-582)."""
    B, N, _ = x.shape
    s = int(math.isqrt(N))
    h = _gelu(_lin(x, sd, p + ".fc1"))
    v1 = cfg.family == "qavit_v1"
    if not v1:
        h = _ln(h, sd, p + ".dwconv_norm")
    Ch = h.shape[-1]
    img = h.transpose(1, 2).reshape(B, Ch, s, s)
    bias = sd.get(p + ".dwconv.dwconv.bias") if cfg.dwconv_bias else None
    img = F.conv2d(img, sd[p + ".dwconv.dwconv.weight"], bias, padding=1, groups=Ch)
    if not v1:
        img = img * sd[p + ".dwconv.scale"]
    h = img.flatten(2).transpose(1, 2)
    if not v1:
        h = _ln(h, sd, p + ".post_dwconv_norm")
    h = _drop(_lin(h, sd, p + ".fc2"), masks, "ffn")                    # H:710 / QAViT.py:582
    if not v1:
        h = h * sd[p + ".gamma"]
    return h


def quad_block(x: Tensor, sd, p: str, cfg: OracleConfig, bank: Bank, train: bool, masks: Optional[dict] = None) -> Tensor:
    """QuadAttentionBlock.forward (H:1071-1085).  ``masks`` (train mode only): explicit keep-scale tensors
    for the block's dropout / DropPath sites -- att_{swa,msda,cga,cross} (attention probabilities),
    proj_{...} (branch outputs), b1 / b2 (BottleneckMLP, H:654-656), ffn (H:710), path1 / path2
    ([B, 1, 1], H:1082-1083); None = dropout 0."""
    xn = _ln(x, sd, p + ".norm1")
    b0 = _lin(_ln(swa(xn, sd, p + ".swa", cfg, bank, train, masks), sd, p + ".norm_swa"), sd, p + ".compress_swa")
    b1 = _lin(_ln(msda(xn, sd, p + ".msda", cfg, bank, train, masks), sd, p + ".norm_msda"), sd, p + ".compress_msda")
    b2 = _lin(_ln(cga(xn, sd, p + ".cga", cfg, bank, train, masks), sd, p + ".norm_cga"), sd, p + ".compress_cga")
    b3 = _lin(_ln(cross(xn, sd, p + ".cross_attn", cfg, bank, masks), sd, p + ".norm_cross"), sd, p + ".compress_cross")
    a = torch.softmax(sd[p + ".fusion.fusion_weights"], 0)                       # H:637-640
    f = torch.cat([b0 * a[0], b1 * a[1], b2 * a[2], b3 * a[3]], -1)
    h = _drop(_gelu(_lin(f, sd, p + ".bottleneck_mlp.fc1")), masks, "b1")
    m = _drop(_lin(h, sd, p + ".bottleneck_mlp.fc2"), masks, "b2")
    x = x + _drop(m, masks, "path1")
    return x + _drop(ccf_ffn(_ln(x, sd, p + ".norm2"), sd, p + ".ccf_ffn", cfg, masks), masks, "path2")


def token_learner(x: Tensor, sd, p: str) -> Tensor:
    """TokenLearner.forward (H:985-1002)."""
    s = torch.softmax(_lin(_ln(x, sd, p + ".attention.0"), sd, p + ".attention.1"), dim=1)
    return s.transpose(1, 2) @ x


def token_upmix(xc: Tensor, sd, p: str) -> Tensor:
    """TokenUpMix.forward (H:1016-1031)."""
    W, b = sd[p + ".upsample_attn.weight"], sd[p + ".upsample_attn.bias"]
    up = torch.einsum("nm,bmc->bnc", W, xc) + b[None, :, None]
    return _ln(up, sd, p + ".norm")


def wrapped_block(x: Tensor, sd, p: str, cfg: OracleConfig, bank: Bank, train: bool, masks: Optional[dict] = None) -> Tensor:
    """QuadBlockWithTokenLearner.forward (H:1104-1123)."""
    if cfg.use_token_learner:
        xc = token_learner(x, sd, p + ".token_learner")
        xc = quad_block(xc, sd, p + ".quad_block", cfg, bank, train, masks)
        return token_upmix(xc, sd, p + ".token_upmix")
    return quad_block(x, sd, p + ".quad_block", cfg, bank, train, masks)


def patch_embed(x: Tensor, sd, cfg: OracleConfig) -> Tensor:
    """PatchEmbed.forward (H:1136-1138) + pos_embed (H:1250).  stride == kernel, so the
    conv is a GEMM over [B*N, 3*p*p] patch rows."""
    B, Cin, S, _ = x.shape
    p = cfg.patch_size
    n = S // p
    rows = x.view(B, Cin, n, p, n, p).permute(0, 2, 4, 1, 3, 5).reshape(B, n * n, Cin * p * p)
    W = sd["patch_embed.proj.weight"].reshape(cfg.embed_dim, -1)
    t = rows @ W.t() + sd["patch_embed.proj.bias"]
    return _ln(t, sd, "patch_embed.norm") + sd["pos_embed"]


# --------------------------------------------------------------------------- HQAViT lateral path
def _bn(x: Tensor, sd, p: str, train: bool, new_stats: Optional[dict]) -> Tensor:
    """nn.BatchNorm2d (H:753): batch stats in train mode, running stats in eval."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    if train:
        mu = x.mean((0, 2, 3))
        var = x.var((0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                new_stats[p + ".running_mean"] = 0.9 * sd[p + ".running_mean"] + 0.1 * mu
                new_stats[p + ".running_var"] = 0.9 * sd[p + ".running_var"] + 0.1 * var * n / (n - 1)
                new_stats[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    else:
        mu, var = sd[p + ".running_mean"], sd[p + ".running_var"]
    xh = (x - mu[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + 1e-5)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def _conv(x, sd, p, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding, groups=groups)


def convnext(x: Tensor, sd, p: str) -> Tensor:
    """ConvNeXtBlock.forward (H:729-739); LN eps 1e-6 (H:723)."""
    C = x.shape[1]
    h = _conv(x, sd, p + ".dwconv", padding=3, groups=C).permute(0, 2, 3, 1)
    h = _ln(h, sd, p + ".norm", eps=1e-6)
    h = _lin(_gelu(_lin(h, sd, p + ".pwconv1")), sd, p + ".pwconv2")
    return x + h.permute(0, 3, 1, 2)


def cnn_stem(x: Tensor, sd, train: bool, new_stats) -> Tuple[Tensor, Tensor, Tensor]:
    """CNNStemModel.forward (H:779-793)."""
    p = "cnn_stem"
    h = _gelu(_bn(_conv(x, sd, p + ".stem.0", 2, 1), sd, p + ".stem.1", train, new_stats))
    h = _gelu(_bn(_conv(h, sd, p + ".stage1.0", 2, 1), sd, p + ".stage1.1", train, new_stats))
    f2 = convnext(h, sd, p + ".stage1.3")
    f3 = convnext(_bn(_conv(f2, sd, p + ".stage2.0"), sd, p + ".stage2.1", train, new_stats), sd, p + ".stage2.2")
    f4 = convnext(_bn(_conv(f3, sd, p + ".stage3.0"), sd, p + ".stage3.1", train, new_stats), sd, p + ".stage3.2")
    return f2, f3, f4


def convnext_v2(x: Tensor, sd, p: str, keep: Optional[Tensor] = None) -> Tensor:
    """ConvNeXtBlock.forward of HQAViTv2_CIFAR100.py:735-751: LayerScale gamma after pwconv2, then DropPath
    (``keep``: per-image keep scale [B], None = identity)."""
    C = x.shape[1]
    h = _conv(x, sd, p + ".dwconv", padding=3, groups=C).permute(0, 2, 3, 1)
    h = _ln(h, sd, p + ".norm", eps=1e-6)
    h = _lin(_gelu(_lin(h, sd, p + ".pwconv1")), sd, p + ".pwconv2")
    h = (sd[p + ".gamma"] * h).permute(0, 3, 1, 2)
    if keep is not None:
        h = h * keep.view(-1, 1, 1, 1)
    return x + h


def _sln(x: Tensor, sd, p: str) -> Tensor:
    """nn.LayerNorm([C, H, W], eps=1e-6) on an NCHW map (V:765, 776, 790)."""
    return F.layer_norm(x, x.shape[1:], sd[p + ".weight"], sd[p + ".bias"], 1e-6)


V2_STEM_BLOCKS = ("stage2.0", "stage2.1", "stage3.0", "stage3.1", "stage3.2", "stage4.0", "stage4.1")


def cnn_stem_v2(x: Tensor, sd, keeps: Optional[dict] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """CNNStemModel.forward of HQAViTv2_CIFAR100.py:811-829.  ``keeps``: {block name: DropPath keep scale [B]}."""
    p = "cnn_stem"
    k = keeps or {}
    h = _sln(_conv(x, sd, p + ".stem.0", stride=4), sd, p + ".stem.1")
    for b in V2_STEM_BLOCKS[0:2]:
        h = convnext_v2(h, sd, f"{p}.{b}", k.get(b))
    f2 = h
    h = _conv(_sln(f2, sd, p + ".downsample2.0"), sd, p + ".downsample2.1")
    for b in V2_STEM_BLOCKS[2:5]:
        h = convnext_v2(h, sd, f"{p}.{b}", k.get(b))
    f3 = h
    h = _conv(_sln(f3, sd, p + ".downsample3.0"), sd, p + ".downsample3.1")
    for b in V2_STEM_BLOCKS[5:7]:
        h = convnext_v2(h, sd, f"{p}.{b}", k.get(b))
    return f2, f3, h


def lmfa(feat: Tensor, sd, p: str, target_hw: int) -> Tensor:
    """LMFAdapter.forward (H:819-849)."""
    C = feat.shape[1]
    f1 = _conv(feat, sd, p + ".dwconv_3x3", padding=1, groups=C)
    f2 = _conv(feat, sd, p + ".dwconv_5x5", padding=2, groups=C)
    fp = _conv(torch.cat([f1, f2, feat], 1), sd, p + ".proj")
    if fp.shape[2] != target_hw or fp.shape[3] != target_hw:
        fp = F.interpolate(fp, size=(target_hw, target_hw), mode="bilinear", align_corners=False)
    return _gelu(_ln(fp.flatten(2).transpose(1, 2), sd, p + ".norm"))


def rrcv(A: Tensor, sd, p: str, cfg: OracleConfig, hw: int) -> Tensor:
    """RRCV.forward (H:880-907)."""
    B, N, C = A.shape
    r = _conv(A.permute(0, 2, 1).reshape(B, C, hw, hw), sd, p + ".reverse_proj")
    for i in range(cfg.rrcv_num_blocks):      # HQAViTv2's RRCV uses its LayerScale block (drop_path 0), V:895
        r = convnext_v2(r, sd, f"{p}.blocks.{i}") if cfg.stem == "v2" else convnext(r, sd, f"{p}.blocks.{i}")
    r = _conv(r, sd, p + ".reembed_proj").flatten(2).transpose(1, 2)
    return A + sd[p + ".beta"] * _ln(r, sd, p + ".norm")


def split_fusion(T: Tensor, R: Tensor, sd, p: str, keep: Optional[Tensor] = None) -> Tensor:
    """SplitFusion.forward (H:941-965); ``keep``: keep-scale tensor [B, N, d] of the hard-coded Dropout(0.1) after the GELU of
    cat_mlp (H:930), None = p 0."""
    gate = torch.sigmoid(_lin(_ln(T + R, sd, p + ".gate_norm"), sd, p + ".gate_fc"))
    t_add = T + gate * R
    m = _gelu(_ln(_lin(torch.cat([T, R], -1), sd, p + ".cat_mlp.0"), sd, p + ".cat_mlp.1"))
    t_cat = T + (m if keep is None else m * keep)
    w = torch.softmax(sd[p + ".fusion_weights"], 0)
    return _ln(w[0] * t_add + w[1] * t_cat, sd, p + ".final_norm")


# --------------------------------------------------------------------------- whole models
def random_masks(cfg: OracleConfig, B: int, Nt: int, p: float, p_path: float,
                 generator: Optional[torch.Generator] = None) -> Dict[str, Tensor]:
    """Bernoulli keep-scale tensors (0 or 1/(1-p)) for every dropout / DropPath site of one
    quad block, drawn the way nn.Dropout does (H:256-264, 416-465, 648-656, 697-710)."""
    d, H, G, kb = cfg.embed_dim, cfg.num_heads, cfg.num_channel_groups, cfg.global_bank_size
    w = cfg.window_size
    nW = Nt // (w * w)
    nkv = cfg.linformer_k + kb

    def keep(shape, q):
        return (torch.rand(shape, generator=generator) >= q).float() / (1.0 - q)

    m: Dict[str, Tensor] = {}
    if p > 0:
        m["att_swa"] = keep((B * nW, H, w * w, nkv), p)
        m["att_msda"] = keep((B, H, Nt, nkv), p)
        m["att_cga"] = keep((B * G, H, Nt, Nt + kb), p)
        m["att_cross"] = keep((B, H, Nt, kb), p)
        for n in ("swa", "msda", "cga", "cross"):
            m["proj_" + n] = keep((B, Nt, d), p)
        m["b1"] = keep((B, Nt, d // cfg.bottleneck_ratio), p)
        m["b2"] = keep((B, Nt, d), p)
        m["ffn"] = keep((B, Nt, d), p)
    if p_path > 0:
        m["path1"] = keep((B, 1, 1), p_path)
        m["path2"] = keep((B, 1, 1), p_path)
    return m


def forward(sd: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, train: bool = False,
            new_state: Optional[dict] = None, mask_fn=None) -> Tensor:
    """QAViT.forward (QAViT.py:689-699) / HQAViT.forward (H:1226-1277) -> logits[B, classes].

    ``new_state`` (train mode) receives the mutated non-autograd state: global_k/global_v,
    update_count and BatchNorm running statistics.  ``mask_fn(block_index, B, tokens)`` (train
    mode, optional) returns the dropout masks of a block (see quad_block); ``mask_fn("pos", B, N)``
    the keep-scale tensor of pos_drop (H:1251); ``mask_fn("fuse2" | "fuse3" | "fuse4", B, N)`` that of
    SplitFusion's Dropout (H:930) or None; ``mask_fn("stem", B, 0)`` (HQAViTv2 stem) a dict
    {block name: DropPath keep scale [B]} or None."""
    bank = Bank(sd, cfg)
    blk_no = [0]

    def masks_for(T):
        if mask_fn is None or not train:
            return None
        nt = cfg.num_learned_tokens if (cfg.family == "hqavit" and cfg.use_token_learner) else T.shape[1]
        blk_no[0] += 1
        return mask_fn(blk_no[0] - 1, T.shape[0], nt)

    def pos_drop(T):
        if mask_fn is None or not train:
            return T
        pm = mask_fn("pos", T.shape[0], T.shape[1])
        return T if pm is None else T * pm

    if cfg.family == "hqavit":
        hw = (cfg.built_img_size or cfg.img_size) // cfg.patch_size      # LMFAdapter.target_hw = the grid the model was built for
        if cfg.stem == "v2":
            keeps = mask_fn("stem", x.shape[0], 0) if (mask_fn is not None and train) else None
            f2, f3, f4 = cnn_stem_v2(x, sd, keeps)
        else:
            f2, f3, f4 = cnn_stem(x, sd, train, new_state)
        R = [rrcv(lmfa(f, sd, f"lmfa{i}", hw), sd, f"rrcv{i}", cfg, hw) for i, f in ((2, f2), (3, f3), (4, f4))]
        T = pos_drop(patch_embed(x, sd, cfg))
        for st, nblk in enumerate(cfg.stage_depths, start=1):
            if st >= 2:
                fk = mask_fn(f"fuse{st}", T.shape[0], T.shape[1]) if (mask_fn is not None and train) else None
                T = split_fusion(T, R[st - 2], sd, f"fuse{st}", fk)
            for i in range(nblk):
                T = wrapped_block(T, sd, f"stage{st}_blocks.{i}", cfg, bank, train, masks_for(T))
    else:
        T = pos_drop(patch_embed(x, sd, cfg))
        for i in range(cfg.depth):
            T = quad_block(T, sd, f"blocks.{i}", cfg, bank, train, masks_for(T))
    T = _ln(T, sd, "norm").mean(1)
    logits = _lin(T, sd, "head")
    if new_state is not None and train:
        k, v = bank.read()
        new_state["global_bank.global_k"] = k.detach().clone()
        new_state["global_bank.global_v"] = v.detach().clone()
        if not bank.v1:
            new_state["global_bank.update_count"] = torch.tensor(bank.count)
    return logits


def cross_entropy(logits: Tensor, y: Tensor, label_smoothing: float = 0.0,
                  y_b: Optional[Tensor] = None, lam: float = 1.0) -> Tensor:
    """nn.CrossEntropyLoss(label_smoothing) mean-reduced (H:1373) and the mixup/cutmix
    two-target form lam*CE(a) + (1-lam)*CE(b) (H:1404-1408)."""
    lp = logits - torch.logsumexp(logits, -1, keepdim=True)

    def one(t):
        nll = -lp.gather(1, t[:, None]).squeeze(1)
        return ((1 - label_smoothing) * nll + label_smoothing * (-lp.mean(-1))).mean()

    if y_b is None:
        return one(y)
    return lam * one(y) + (1 - lam) * one(y_b)


def clip_grads_(grads: Dict[str, Optional[Tensor]], max_norm: float = 0.5,
                per_param_names=("cnn_stem", "dwconv"), per_param_max: float = 0.1) -> float:
    """H:1416-1418 per-parameter clip then H:1432 global clip_grad_norm_.  Returns the
    global norm measured before the global clip (what clip_grad_norm_ returns)."""
    for n, g in grads.items():
        if g is not None and any(s in n for s in per_param_names):
            g.mul_(torch.clamp(per_param_max / (g.norm() + 1e-6), max=1.0))
    tot = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values() if g is not None)).float()
    coef = torch.clamp(max_norm / (tot + 1e-6), max=1.0)
    for g in grads.values():
        if g is not None:
            g.mul_(coef)
    return float(tot)


def adamw_step_(params: Dict[str, Tensor], grads: Dict[str, Optional[Tensor]], state: dict, lr: float,
                beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, wd: float = 0.06) -> None:
    """torch.optim.AdamW single-tensor rule (H:1566-1571).  Params whose grad is
    None are skipped entirely (no weight decay either)."""
    for n, p in params.items():
        g = grads.get(n)
        if g is None:
            continue
        st = state.setdefault(n, {"t": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["t"] += 1
        t = st["t"]
        p.mul_(1 - lr * wd)
        st["m"].mul_(beta1).add_(g, alpha=1 - beta1)
        st["v"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (st["v"].sqrt() / math.sqrt(1 - beta2 ** t)).add_(eps)
        p.addcdiv_(st["m"], denom, value=-lr / (1 - beta1 ** t))


# --------------------------------------------------------------------------- schema + synthetic weights
def state_schema(cfg: OracleConfig) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """Ordered {key: (shape, kind)} of the *unique* state_dict entries.
    The reference additionally aliases global_bank.* under every branch; see
    ``with_bank_aliases``.  kind in: lin_w lin_b ln_w ln_b conv_w conv_b bank lf small_pos
    scale gamma beta fw2 fw4 pos bn_mean bn_var count."""
    d, Kb, H = cfg.embed_dim, cfg.global_bank_size, cfg.num_heads
    G = cfg.num_channel_groups
    n_tok = (cfg.img_size // cfg.patch_size) ** 2                          # pos_embed rows (resized to the input grid)
    n_up = ((cfg.built_img_size or cfg.img_size) // cfg.patch_size) ** 2   # TokenUpMix output tokens (fixed at construction)
    p = cfg.patch_size
    S: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def lin(k, o, i):
        S[k + ".weight"] = ((o, i), "lin_w")
        S[k + ".bias"] = ((o,), "lin_b")

    def ln(k, c):
        S[k + ".weight"] = ((c,), "ln_w")
        S[k + ".bias"] = ((c,), "ln_b")

    def conv(k, o, i, ks, bias=True):
        S[k + ".weight"] = ((o, i, ks, ks), "conv_w")
        if bias:
            S[k + ".bias"] = ((o,), "conv_b")

    def bn(k, c):
        ln(k, c)
        S[k + ".running_mean"] = ((c,), "bn_mean")
        S[k + ".running_var"] = ((c,), "bn_var")
        S[k + ".num_batches_tracked"] = ((), "count")

    def cnx(k, c):
        conv(k + ".dwconv", c, 1, 7)
        ln(k + ".norm", c)
        lin(k + ".pwconv1", 4 * c, c)
        lin(k + ".pwconv2", c, 4 * c)

    def block(P, ntok_blk):
        ws2 = cfg.window_size ** 2
        ln(P + ".norm1", d)
        lin(P + ".swa.qkv", 3 * d, d)
        S[P + ".swa.linformer.E_k"] = ((ws2, cfg.linformer_k), "lf")
        S[P + ".swa.linformer.E_v"] = ((ws2, cfg.linformer_k), "lf")
        lin(P + ".swa.proj", d, d)
        ln(P + ".swa.norm", d)
        lin(P + ".msda.qkv", 3 * d, d)
        S[P + ".msda.linformer.E_k"] = ((cfg.msda_seq_len, cfg.linformer_k), "lf")
        S[P + ".msda.linformer.E_v"] = ((cfg.msda_seq_len, cfg.linformer_k), "lf")
        lin(P + ".msda.proj", d, d)
        ln(P + ".msda.norm", d)
        cpg = (d // 2) // G
        for n in ("q_proj", "k_proj", "v_proj"):
            lin(f"{P}.cga.{n}", cpg, d // G)
        lin(P + ".cga.bank_k_proj", cpg, d)
        lin(P + ".cga.bank_v_proj", cpg, d)
        lin(P + ".cga.proj", d, d // 2)
        ln(P + ".cga.norm", d)
        for n in ("q_proj", "k_proj", "v_proj", "proj"):
            lin(f"{P}.cross_attn.{n}", d, d)
        for n in ("swa", "msda", "cga", "cross"):
            ln(f"{P}.norm_{n}", d)
        for n in ("swa", "msda", "cga", "cross"):
            lin(f"{P}.compress_{n}", d // cfg.compress_ratio, d)
        S[P + ".fusion.fusion_weights"] = ((4,), "fw4")
        lin(P + ".bottleneck_mlp.fc1", d // cfg.bottleneck_ratio, d)
        lin(P + ".bottleneck_mlp.fc2", d, d // cfg.bottleneck_ratio)
        ln(P + ".norm2", d)
        hid = int(d * cfg.mlp_ratio)
        if cfg.family != "qavit_v1":
            S[P + ".ccf_ffn.gamma"] = ((1,), "gamma")
        lin(P + ".ccf_ffn.fc1", hid, d)
        if cfg.family != "qavit_v1":
            ln(P + ".ccf_ffn.dwconv_norm", hid)
            S[P + ".ccf_ffn.dwconv.scale"] = ((1, hid, 1, 1), "scale")
        conv(P + ".ccf_ffn.dwconv.dwconv", hid, 1, 3, bias=cfg.dwconv_bias)
        if cfg.family != "qavit_v1":
            ln(P + ".ccf_ffn.post_dwconv_norm", hid)
        lin(P + ".ccf_ffn.fc2", d, hid)

    S["pos_embed"] = ((1, n_tok, d), "pos")
    conv("patch_embed.proj", d, cfg.in_channels, p)
    ln("patch_embed.norm", d)
    S["global_bank.global_k"] = ((1, Kb, d), "bank")
    S["global_bank.global_v"] = ((1, Kb, d), "bank")
    ln("global_bank.write_norm", d)
    lin("global_bank.write_compression", d, d)
    lin("global_bank.write_gate", Kb, d)
    if cfg.family != "qavit_v1":
        S["global_bank.update_count"] = ((), "count")
    if cfg.family == "hqavit" and cfg.stem == "v2":
        g = (cfg.built_img_size or cfg.img_size) // 4

        def sln(k, c):
            S[k + ".weight"] = ((c, g, g), "ln_w")
            S[k + ".bias"] = ((c, g, g), "ln_b")

        def cnx2(k, c):
            S[k + ".gamma"] = ((c,), "ls_gamma")
            cnx(k, c)

        conv("cnn_stem.stem.0", cfg.cnn_c2, cfg.in_channels, 4)
        sln("cnn_stem.stem.1", cfg.cnn_c2)
        cnx2("cnn_stem.stage2.0", cfg.cnn_c2)
        cnx2("cnn_stem.stage2.1", cfg.cnn_c2)
        sln("cnn_stem.downsample2.0", cfg.cnn_c2)
        conv("cnn_stem.downsample2.1", cfg.cnn_c3, cfg.cnn_c2, 1)
        for i in range(3):
            cnx2(f"cnn_stem.stage3.{i}", cfg.cnn_c3)
        sln("cnn_stem.downsample3.0", cfg.cnn_c3)
        conv("cnn_stem.downsample3.1", cfg.cnn_c4, cfg.cnn_c3, 1)
        for i in range(2):
            cnx2(f"cnn_stem.stage4.{i}", cfg.cnn_c4)
    if cfg.family == "hqavit" and cfg.stem != "v2":
        conv("cnn_stem.stem.0", 32, cfg.in_channels, 3)
        bn("cnn_stem.stem.1", 32)
        conv("cnn_stem.stage1.0", cfg.cnn_c2, 32, 3)
        bn("cnn_stem.stage1.1", cfg.cnn_c2)
        cnx("cnn_stem.stage1.3", cfg.cnn_c2)
        conv("cnn_stem.stage2.0", cfg.cnn_c3, cfg.cnn_c2, 1)
        bn("cnn_stem.stage2.1", cfg.cnn_c3)
        cnx("cnn_stem.stage2.2", cfg.cnn_c3)
        conv("cnn_stem.stage3.0", cfg.cnn_c4, cfg.cnn_c3, 1)
        bn("cnn_stem.stage3.1", cfg.cnn_c4)
        cnx("cnn_stem.stage3.2", cfg.cnn_c4)
    if cfg.family == "hqavit":
        for i, c in ((2, cfg.cnn_c2), (3, cfg.cnn_c3), (4, cfg.cnn_c4)):
            conv(f"lmfa{i}.dwconv_3x3", c, 1, 3)
            conv(f"lmfa{i}.dwconv_5x5", c, 1, 5)
            conv(f"lmfa{i}.proj", d, 3 * c, 1)
            ln(f"lmfa{i}.norm", d)
        for i in (2, 3, 4):
            S[f"rrcv{i}.beta"] = ((), "beta")
            conv(f"rrcv{i}.reverse_proj", cfg.rrcv_channels, d, 1)
            for j in range(cfg.rrcv_num_blocks):
                if cfg.stem == "v2":
                    S[f"rrcv{i}.blocks.{j}.gamma"] = ((cfg.rrcv_channels,), "ls_gamma")
                cnx(f"rrcv{i}.blocks.{j}", cfg.rrcv_channels)
            conv(f"rrcv{i}.reembed_proj", d, cfg.rrcv_channels, 1)
            ln(f"rrcv{i}.norm", d)
        for i in (2, 3, 4):
            S[f"fuse{i}.fusion_weights"] = ((2,), "fw2")
            ln(f"fuse{i}.gate_norm", d)
            lin(f"fuse{i}.gate_fc", d, d)
            lin(f"fuse{i}.cat_mlp.0", d, 2 * d)
            ln(f"fuse{i}.cat_mlp.1", d)
            ln(f"fuse{i}.final_norm", d)
        M = cfg.num_learned_tokens
        for st, nblk in enumerate(cfg.stage_depths, start=1):
            for i in range(nblk):
                P = f"stage{st}_blocks.{i}"
                if cfg.use_token_learner:
                    ln(P + ".token_learner.attention.0", d)
                    lin(P + ".token_learner.attention.1", M, d)
                    lin(P + ".token_upmix.upsample_attn", n_up, M)
                    ln(P + ".token_upmix.norm", d)
                block(P + ".quad_block", M if cfg.use_token_learner else n_tok)
    else:
        for i in range(cfg.depth):
            block(f"blocks.{i}", n_tok)
    ln("norm", d)
    lin("head", cfg.num_classes, d)
    return S


_KIND_RECIPE = {
    # kind: (mean, std) -- larger than the reference init on purpose: a parity test with
    # near-uniform softmaxes and near-zero biases would hide indexing bugs.
    "lin_b": (0.0, 0.02), "ln_w": (1.0, 0.1), "ln_b": (0.0, 0.05), "conv_b": (0.0, 0.02),
    "bank": (0.0, 0.1), "lf": (0.0, 0.15), "scale": (0.1, 0.02), "gamma": (0.5, 0.0),
    "beta": (0.1, 0.0), "fw2": (0.5, 0.3), "fw4": (1.0, 0.5), "pos": (0.0, 0.1),
    "bn_mean": (0.0, 0.1),
    "ls_gamma": (0.3, 0.1),      # LayerScale: the reference's 1e-6 init would hide every bug behind it
}


def synthetic_state(cfg: OracleConfig, seed: int = 1234) -> Dict[str, Tensor]:
    """Deterministic synthetic weights, a pure function of (schema order, seed): one CPU
    generator, keys drawn in schema order.  Used for golden vectors and GPU parity tests so
    the fixtures need not carry 26 MB of weights."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for k, (shape, kind) in state_schema(cfg).items():
        if kind == "count":
            out[k] = torch.tensor(0, dtype=torch.int64)
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if kind == "lin_w":
            t = r * (1.0 / math.sqrt(shape[1]))
        elif kind == "conv_w":
            fan_in = shape[1] * shape[2] * shape[3]
            t = r * (1.0 / math.sqrt(fan_in))
        elif kind == "bn_var":
            t = 1.0 + 0.2 * r.abs()
        else:
            m, s = _KIND_RECIPE[kind]
            t = m + s * r
        out[k] = t
    return out


BRANCHES = ("swa", "msda", "cga", "cross_attn")
BANK_KEYS = ("global_k", "global_v", "write_norm.weight", "write_norm.bias", "write_compression.weight",
             "write_compression.bias", "write_gate.weight", "write_gate.bias", "update_count")


def block_prefixes(cfg: OracleConfig):
    if cfg.family == "hqavit":
        return [f"stage{st}_blocks.{i}.quad_block" for st, n in enumerate(cfg.stage_depths, 1) for i in range(n)]
    return [f"blocks.{i}" for i in range(cfg.depth)]


def with_bank_aliases(sd: Dict[str, Tensor], cfg: OracleConfig) -> Dict[str, Tensor]:
    """Add the aliased ``<block>.<branch>.global_bank.*`` entries the reference's
    state_dict carries (one GlobalTokenBank registered under every branch, H:407,476,539,602)."""
    out = dict(sd)
    for P in block_prefixes(cfg):
        for br in BRANCHES:
            for bk in BANK_KEYS:
                if "global_bank." + bk in sd:
                    out[f"{P}.{br}.global_bank.{bk}"] = sd["global_bank." + bk]
    return out


def trainable_keys(cfg: OracleConfig):
    """Float parameters (not buffers) in schema order."""
    return [k for k, (_, kind) in state_schema(cfg).items() if kind not in ("count", "bn_mean", "bn_var")]


def loss_and_grads(sd: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, y: Tensor, label_smoothing: float = 0.0,
                   train: bool = True, mask_fn=None):
    """One training forward+backward.  Returns (logits, loss, grads{key: tensor|None}, new_state).
    ``mask_fn``: see forward()."""
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in trainable_keys(cfg)}
    full = dict(sd)
    full.update(leaves)
    new_state: dict = {}
    logits = forward(full, cfg, x, train=train, new_state=new_state, mask_fn=mask_fn)
    loss = cross_entropy(logits, y, label_smoothing)
    gl = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = {k: g for k, g in zip(leaves.keys(), gl)}
    return logits.detach(), loss.detach(), grads, new_state
