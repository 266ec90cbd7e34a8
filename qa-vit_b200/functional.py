"""torch.autograd.Functions over the C ABI: each one stands in for the autograd graph of one reference module.

All tensors stay on the device; the C side runs on torch's current stream, allocates nothing and never
synchronises, so the Functions are safe under autocast and CUDA-graph capture."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import BlockCfg, LateralCfg, QP_COUNT, SplitFusionCfg, check, lib

_scratch = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _scratch_buf(device, nbytes: int) -> torch.Tensor:
    """One reusable scratch allocation per (device, stream): temporaries of a native call, reused in stream order.
    Keyed by stream because HQAViT runs its lateral path on a side stream concurrently with the stage-1 blocks."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"qavit_b200: {what} must be a CUDA tensor -- there is no CPU path in this package")


def resolve_dtype(precision: str) -> int:
    """'fp32' | 'bf16' | 'auto' (bf16 under torch.autocast(bfloat16), like the reference train loop H:1402)."""
    if precision == "auto":
        return 1 if (torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16) else 0
    return {"fp32": 0, "bf16": 1}[precision]


_grad_ready_callbacks = []


def register_grad_ready_callback(fn):
    """fn(list_of_parameters) is called at the end of every native backward with the parameters whose gradients were
    accumulated IN PLACE into an already-attached .grad (autograd's post-accumulate hooks do not fire for those)."""
    _grad_ready_callbacks.append(fn)
    return fn


def unregister_grad_ready_callback(fn):
    if fn in _grad_ready_callbacks:
        _grad_ready_callbacks.remove(fn)


def _inplace_target(t: torch.Tensor) -> Optional[torch.Tensor]:
    """The attached gradient buffer of a leaf parameter, if the kernels can accumulate straight into it."""
    g = getattr(t, "grad", None)
    if g is not None and t.is_leaf and g.dtype == torch.float32 and g.is_contiguous() and g.device == t.device:
        return g
    return None


def _notify(params):
    if params and _grad_ready_callbacks:
        for fn in list(_grad_ready_callbacks):
            fn(params)


def _grad_out(t: torch.Tensor, direct: list):
    """(buffer the kernel accumulates into, value to return to autograd).  With .grad attached (FusedAdamW's flat
    buffer) the kernel adds in place and autograd gets None: no add kernel, no copy."""
    g = _inplace_target(t)
    if g is not None:
        direct.append(t)
        return g, None
    z = torch.zeros_like(t, dtype=torch.float32)
    return z, z


class BlockMeta:
    """Static description of one block call: the cfg struct and which tensor sits at which QP_* index."""

    def __init__(self, cfg: BlockCfg, index: Sequence[int], no_grad_idx: Sequence[int], update_count: Optional[torch.Tensor]):
        self.cfg = cfg
        self.index = list(index)               # QP index of the i-th tensor argument
        self.no_grad = set(no_grad_idx)        # QP indices that never receive a gradient (write_*, branch .norm)
        self.update_count = update_count


class QuadBlockFn(torch.autograd.Function):
    """QuadAttentionBlock.forward (+ TokenLearner / TokenUpMix wrapper) -- HQAViT_CIFAR100.py:1071-1123."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, meta: BlockMeta, *tensors: torch.Tensor):
        _require_cuda(x, "block input")
        x = x.detach().float().contiguous()
        cfg = meta.cfg
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_block_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        saved = torch.empty(saved_b.value, dtype=torch.uint8, device=x.device)
        scratch = _scratch_buf(x.device, scratch_b.value)
        params = (C.c_void_p * QP_COUNT)()
        for qi, t in zip(meta.index, tensors):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("qavit_b200: parameters must be contiguous fp32 tensors")
            params[qi] = t.data_ptr()
        n_out = cfg.tokens_out if (cfg.token_learner and cfg.tokens_out > 0) else x.shape[1]
        out = torch.empty(x.shape[0], n_out, x.shape[2], dtype=torch.float32, device=x.device)
        rng = rng_state(x.device) if (cfg.train and (cfg.dropout > 0 or cfg.drop_path > 0)) else None
        check(lib.qavit_block_forward(C.byref(cfg), params, _ptr(meta.update_count), _ptr(rng), x.data_ptr(), out.data_ptr(),
                                      saved.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.meta, ctx.saved_buf, ctx.x, ctx.params_arr = meta, saved, x, params
        ctx.tensors = tensors
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        meta, x = ctx.meta, ctx.x
        cfg = meta.cfg
        dout = dout.float().contiguous()
        # gradients accumulate straight into an attached .grad (flat optimizer buffer) when there is one; the rest
        # share one zeroed flat allocation whose views are handed back to autograd.
        sizes: List[int] = []
        direct: List[torch.Tensor] = []
        targets: List[Optional[torch.Tensor]] = []
        for qi, t in zip(meta.index, ctx.tensors):
            if qi in meta.no_grad:
                targets.append(None)
                sizes.append(0)
                continue
            g = _inplace_target(t)
            targets.append(g)
            sizes.append(0 if g is not None else (t.numel() + 3) // 4 * 4)
            if g is not None:
                direct.append(t)
        gbuf = torch.zeros(sum(sizes), dtype=torch.float32, device=x.device) if sum(sizes) else None
        grads_arr = (C.c_void_p * QP_COUNT)()
        views: List[Optional[torch.Tensor]] = []
        off = 0
        for (qi, t), n, g in zip(zip(meta.index, ctx.tensors), sizes, targets):
            if g is not None:
                grads_arr[qi] = g.data_ptr()
                views.append(None)
            elif n == 0:
                views.append(None)
            else:
                v = gbuf[off:off + t.numel()].view(t.shape)
                grads_arr[qi] = v.data_ptr()
                views.append(v)
                off += n
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_block_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        scratch = _scratch_buf(x.device, scratch_b.value)
        dx = torch.empty_like(x)
        check(lib.qavit_block_backward(C.byref(cfg), ctx.params_arr, grads_arr, x.data_ptr(), dout.data_ptr(), dx.data_ptr(),
                                       ctx.saved_buf.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.saved_buf = None
        _notify(direct)
        return (dx, None, *views)


class PatchEmbedFn(torch.autograd.Function):
    """PatchEmbed.forward (+ pos_embed) -- HQAViT_CIFAR100.py:1136-1138, 1250."""

    @staticmethod
    def forward(ctx, img, W, bias, ln_w, ln_b, pos, dtype=0):
        _require_cuda(img, "image batch")
        img = img.detach().float().contiguous()
        B, Cin, S, _ = img.shape
        d, _, p, _ = W.shape
        N = (S // p) ** 2
        pre = torch.empty(B * N, d, dtype=torch.float32, device=img.device)
        stats = torch.empty(B * N, 2, dtype=torch.float32, device=img.device)
        out = torch.empty(B, N, d, dtype=torch.float32, device=img.device)
        scratch = _scratch_buf(img.device, lib.qavit_patch_embed_scratch_bytes(B, Cin, S, p, d)) if dtype == 1 else None
        check(lib.qavit_patch_embed_forward(img.data_ptr(), B, Cin, S, p, d, W.data_ptr(), bias.data_ptr(), ln_w.data_ptr(),
                                            ln_b.data_ptr(), _ptr(pos), pre.data_ptr(), stats.data_ptr(), out.data_ptr(),
                                            dtype, _ptr(scratch), _stream()))
        ctx.dtype = dtype
        ctx.save_for_backward(img, W, ln_w, pre, stats)
        ctx.params = (W, bias, ln_w, ln_b, pos)
        ctx.has_pos = pos is not None
        ctx.pos_shape = None if pos is None else pos.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        img, W, ln_w, pre, stats = ctx.saved_tensors
        B, Cin, S, _ = img.shape
        d, _, p, _ = W.shape
        dout = dout.float().contiguous()
        dev = img.device
        direct = []
        (dW, rW), (db, rb), (dg, rg), (dbeta, rbeta) = (_grad_out(t, direct) for t in ctx.params[:4])
        dpos, rpos = _grad_out(ctx.params[4], direct) if ctx.has_pos else (None, None)
        dpre = torch.empty_like(pre) if ctx.dtype != 1 else None
        scratch = _scratch_buf(dev, lib.qavit_patch_embed_scratch_bytes(B, Cin, S, p, d)) if ctx.dtype == 1 else None
        check(lib.qavit_patch_embed_backward(img.data_ptr(), dout.data_ptr(), B, Cin, S, p, d, pre.data_ptr(), stats.data_ptr(),
                                             ln_w.data_ptr(), _ptr(dpre), dW.data_ptr(), db.data_ptr(), dg.data_ptr(),
                                             dbeta.data_ptr(), _ptr(dpos), ctx.dtype, _ptr(scratch), _stream()))
        _notify(direct)
        return None, rW, rb, rg, rbeta, rpos, None


class HeadFn(torch.autograd.Function):
    """norm -> mean over tokens -> head -- HQAViT_CIFAR100.py:1273-1275."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, W, bias):
        _require_cuda(x, "head input")
        x = x.detach().float().contiguous()
        B, N, d = x.shape
        ncls = W.shape[0]
        stats = torch.empty(B * N, 2, dtype=torch.float32, device=x.device)
        pooled = torch.empty(B, d, dtype=torch.float32, device=x.device)
        logits = torch.empty(B, ncls, dtype=torch.float32, device=x.device)
        check(lib.qavit_head_forward(x.data_ptr(), B, N, d, ln_w.data_ptr(), ln_b.data_ptr(), W.data_ptr(), bias.data_ptr(), ncls,
                                     stats.data_ptr(), pooled.data_ptr(), logits.data_ptr(), _stream()))
        ctx.save_for_backward(x, ln_w, W, stats, pooled)
        ctx.params = (ln_w, ln_b, W, bias)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, ln_w, W, stats, pooled = ctx.saved_tensors
        B, N, d = x.shape
        ncls = W.shape[0]
        dlogits = dlogits.float().contiguous()
        dev = x.device
        dx = torch.empty_like(x)
        direct = []
        (dg, rg), (dbeta, rbeta), (dW, rW), (db, rb) = (_grad_out(t, direct) for t in ctx.params)
        dpooled = torch.empty_like(pooled)
        check(lib.qavit_head_backward(x.data_ptr(), dlogits.data_ptr(), B, N, d, ln_w.data_ptr(), stats.data_ptr(), pooled.data_ptr(),
                                      W.data_ptr(), ncls, dpooled.data_ptr(), dx.data_ptr(), dg.data_ptr(), dbeta.data_ptr(),
                                      dW.data_ptr(), db.data_ptr(), _stream()))
        _notify(direct)
        return dx, rg, rbeta, rW, rb


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last axis (C <= 256) for the modules around the blocks; fp32 output like autocast."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        _require_cuda(x, "LayerNorm input")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.detach().contiguous()
        C_ = x.shape[-1]
        rows = x.numel() // C_
        y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        stats = torch.empty(rows, 2, dtype=torch.float32, device=x.device)
        check(lib.qavit_layer_norm_forward(x.data_ptr(), int(x.dtype == torch.bfloat16), rows, C_, w.data_ptr(), b.data_ptr(),
                                           float(eps), y.data_ptr(), stats.data_ptr(), _stream()))
        ctx.save_for_backward(x, w, stats)
        ctx.params = (w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, stats = ctx.saved_tensors
        C_ = x.shape[-1]
        rows = x.numel() // C_
        dy = dy.float().contiguous()
        dx = torch.empty_like(x)
        direct = []
        (dg, rg), (db, rb) = (_grad_out(t, direct) for t in ctx.params)
        check(lib.qavit_layer_norm_backward(x.data_ptr(), int(x.dtype == torch.bfloat16), dy.data_ptr(), rows, C_, w.data_ptr(),
                                            stats.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), _stream()))
        _notify(direct)
        return dx, rg, rb, None


class DepthwiseConvFn(torch.autograd.Function):
    """Depthwise k x k conv (stride 1, same padding, bias) on a channels-last feature map -- ConvNeXtBlock.dwconv
    (H:722) and LMFAdapter.dwconv_3x3 / 5x5 (H:811-812).  x: logical [B, C, H, W], channels_last strides."""

    @staticmethod
    def forward(ctx, x, w, bias):
        _require_cuda(x, "conv input")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.detach().contiguous(memory_format=torch.channels_last)
        B, C_, H, W = x.shape
        K = w.shape[-1]
        y = torch.empty_like(x, memory_format=torch.channels_last)
        check(lib.qavit_dwconv_forward(x.data_ptr(), int(x.dtype == torch.bfloat16), B, H, W, C_, K, w.data_ptr(), _ptr(bias),
                                       y.data_ptr(), _stream()))
        ctx.save_for_backward(x, w)
        ctx.params = (w, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, C_, H, W = x.shape
        K = w.shape[-1]
        dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x, memory_format=torch.channels_last) if ctx.needs_input_grad[0] else None
        direct = []
        dw, rw = _grad_out(ctx.params[0], direct)
        db, rb = _grad_out(ctx.params[1], direct) if ctx.params[1] is not None else (None, None)
        check(lib.qavit_dwconv_backward(x.data_ptr(), dy.data_ptr(), int(x.dtype == torch.bfloat16), B, H, W, C_, K, w.data_ptr(),
                                        _ptr(dx), dw.data_ptr(), _ptr(db), _stream()))
        _notify(direct)
        return dx, rw, rb


def depthwise_conv(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    """conv(x) for a depthwise, stride-1, same-padded nn.Conv2d with k in (3, 5, 7) through the library's kernels."""
    x = x.to(torch.get_autocast_gpu_dtype()) if torch.is_autocast_enabled() and x.dtype == torch.float32 else x
    return DepthwiseConvFn.apply(x, conv.weight, conv.bias)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b on [M, K] rows (nn.Linear, and 1x1 convolutions on channels-last maps) through the library's GEMMs."""

    @staticmethod
    def forward(ctx, x, W, bias):
        _require_cuda(x, "linear input")
        x = x.detach().contiguous()
        M, K = x.shape
        N = W.shape[0]
        bf = x.dtype == torch.bfloat16
        y = torch.empty(M, N, dtype=x.dtype, device=x.device)
        wb = torch.empty(N * K, dtype=torch.bfloat16, device=x.device) if bf else None
        check(lib.qavit_linear_forward(x.data_ptr(), int(bf), M, K, W.data_ptr(), _ptr(bias), N, y.data_ptr(), _ptr(wb), _stream()))
        ctx.save_for_backward(x, W)
        ctx.params = (W, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        M, K = x.shape
        N = W.shape[0]
        bf = x.dtype == torch.bfloat16
        dy = dy.to(x.dtype).contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        direct = []
        dW, rW = _grad_out(ctx.params[0], direct)
        db, rb = _grad_out(ctx.params[1], direct) if ctx.params[1] is not None else (None, None)
        wbt = torch.empty(N * K, dtype=torch.bfloat16, device=x.device) if bf else None
        check(lib.qavit_linear_backward(x.data_ptr(), dy.data_ptr(), int(bf), M, K, N, W.data_ptr(), _ptr(dx), dW.data_ptr(), _ptr(db),
                                        _ptr(wbt), _stream()))
        _notify(direct)
        return dx, rW, rb


def linear(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """F.linear(x, W, bias) over the last axis (W may be a [N, K, 1, 1] conv weight): bf16 under autocast, else fp32."""
    K = x.shape[-1]
    if torch.is_autocast_enabled():
        x = x.to(torch.get_autocast_gpu_dtype())
    elif x.dtype != torch.float32:
        x = x.float()
    y = LinearFn.apply(x.reshape(-1, K), W.view(W.shape[0], -1) if W.dim() != 2 else W, bias)
    return y.view(*x.shape[:-1], W.shape[0])


def conv1x1(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    """1x1 nn.Conv2d on a (channels-last) [B, C, H, W] map as a GEMM over its [B*H*W, C] rows."""
    B, C_, H, W_ = x.shape
    y = linear(x.permute(0, 2, 3, 1), conv.weight, conv.bias)          # [B, H, W, N]; no copy for channels_last x
    return y.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------ dropout RNG state
_rng_states = {}
_rng_seed = None        # set by manual_seed(); states created later start from it instead of torch.initial_seed()


def rng_state(device) -> torch.Tensor:
    """Device-resident Philox state [seed, offset] (int64) behind every in-kernel dropout: forward calls snapshot it
    into their saved buffer and advance the offset on the device, so CUDA-graph replays draw fresh masks."""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    t = _rng_states.get(device)
    if t is None:
        seed = torch.initial_seed() if _rng_seed is None else _rng_seed
        t = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        _rng_states[device] = t
    return t


def manual_seed(seed: int) -> None:
    """Re-seed the dropout generators of every device (offset back to 0); data-parallel ranks should pass distinct
    seeds (e.g. seed + rank) so that replicas draw independent masks."""
    global _rng_seed
    _rng_seed = int(seed)
    for dev, t in _rng_states.items():
        t.copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))


class DropoutFn(torch.autograd.Function):
    """``nn.Dropout`` on an fp32 tensor (pos_drop, HQAViT_CIFAR100.py:1155 / 1251) with the library's Philox stream."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, p: float):
        _require_cuda(x, "dropout input")
        x = x.float().contiguous()
        if x.numel() % 8:
            raise RuntimeError("qavit_b200: dropout needs a multiple of 8 elements")
        y = torch.empty_like(x)
        snap = torch.empty(2, dtype=torch.int64, device=x.device)
        check(lib.qavit_dropout_forward(x.data_ptr(), y.data_ptr(), x.numel(), float(p), rng_state(x.device).data_ptr(),
                                        snap.data_ptr(), _stream()))
        ctx.snap, ctx.p = snap, float(p)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        dy = dy.float().contiguous()
        dx = torch.empty_like(dy)
        check(lib.qavit_dropout_backward(dy.data_ptr(), dx.data_ptr(), dy.numel(), ctx.p, ctx.snap.data_ptr(), _stream()))
        return dx, None


def dropout(x: torch.Tensor, p: float, training: bool = True) -> torch.Tensor:
    """Drop-in for ``F.dropout(x, p, training)`` on CUDA fp32 tensors (mask regenerated in backward, graph-capture safe)."""
    if not training or p <= 0:
        return x
    return DropoutFn.apply(x, p)


def _param_array(tensors, n):
    arr = (C.c_void_p * n)()
    for i, t in enumerate(tensors):
        if not t.is_contiguous() or t.dtype not in (torch.float32, torch.int64):
            raise RuntimeError("qavit_b200: parameters / buffers must be contiguous fp32 (int64 counters) tensors")
        arr[i] = t.data_ptr()
    return arr


def _grad_arrays(tensors, skip, device):
    """ctypes array of gradient accumulators (+ the values to hand back to autograd): attached .grad buffers are
    accumulated in place, the rest share one zeroed flat allocation."""
    direct, sizes, targets = [], [], []
    for i, t in enumerate(tensors):
        if i in skip or not t.requires_grad:
            targets.append(None)
            sizes.append(0)
            continue
        g = _inplace_target(t)
        targets.append(g)
        sizes.append(0 if g is not None else (t.numel() + 3) // 4 * 4)
        if g is not None:
            direct.append(t)
    gbuf = torch.zeros(sum(sizes), dtype=torch.float32, device=device) if sum(sizes) else None
    arr = (C.c_void_p * len(tensors))()
    views, off = [], 0
    for i, (t, n, g) in enumerate(zip(tensors, sizes, targets)):
        if g is not None:
            arr[i] = g.data_ptr()
            views.append(None)
        elif n == 0:
            views.append(None)
        else:
            v = gbuf[off:off + t.numel()].view(t.shape)
            arr[i] = v.data_ptr()
            views.append(v)
            off += n
    return arr, views, direct


class LateralMeta:
    def __init__(self, cfg: LateralCfg, buffer_idx):
        self.cfg = cfg
        self.buffers = set(buffer_idx)      # indices of BatchNorm running statistics / counters (no gradient)


class LateralFn(torch.autograd.Function):
    """CNNStemModel -> LMFAdapter x3 -> RRCV x3 (HQAViT_CIFAR100.py:1236-1247): image -> (R2, R3, R4)."""

    @staticmethod
    def forward(ctx, img, meta: LateralMeta, *tensors):
        _require_cuda(img, "image batch")
        img = img.detach().float().contiguous()
        cfg = meta.cfg
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_lateral_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        saved = torch.empty(saved_b.value, dtype=torch.uint8, device=img.device)
        scratch = _scratch_buf(img.device, scratch_b.value)
        params = _param_array(tensors, len(tensors))
        N = cfg.grid * cfg.grid
        R = [torch.empty(cfg.batch, N, cfg.dim, dtype=torch.float32, device=img.device) for _ in range(3)]
        check(lib.qavit_lateral_forward(C.byref(cfg), params, img.data_ptr(), R[0].data_ptr(), R[1].data_ptr(), R[2].data_ptr(),
                                        saved.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.meta, ctx.saved_buf, ctx.img, ctx.params_arr, ctx.tensors = meta, saved, img, params, tensors
        return tuple(R)

    @staticmethod
    def backward(ctx, dR2, dR3, dR4):
        meta, img = ctx.meta, ctx.img
        cfg = meta.cfg
        dRs = [None if g is None else g.float().contiguous() for g in (dR2, dR3, dR4)]
        grads_arr, views, direct = _grad_arrays(ctx.tensors, meta.buffers, img.device)
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_lateral_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        scratch = _scratch_buf(img.device, scratch_b.value)
        check(lib.qavit_lateral_backward(C.byref(cfg), ctx.params_arr, grads_arr, img.data_ptr(), _ptr(dRs[0]), _ptr(dRs[1]),
                                         _ptr(dRs[2]), ctx.saved_buf.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.saved_buf = None
        _notify(direct)
        return (None, None, *views)


class LateralState:
    """What the phases of one lateral forward / backward share (qavit_lateral_*_parts): the configuration, the full parameter
    pointer table (the C side indexes it globally), the `saved` buffer and which adapters' backward has run."""

    def __init__(self, cfg: LateralCfg, names, tensors, buffer_idx, img):
        self.cfg, self.names, self.tensors, self.buffers, self.img = cfg, names, tensors, set(buffer_idx), img
        self.params = _param_array(tensors, len(tensors))
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_lateral_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        self.scratch_bytes = scratch_b.value
        self.saved = torch.empty(saved_b.value, dtype=torch.uint8, device=img.device)
        self.bwd_done = [False, False, False]
        # indices of each part's tensors: 0 = cnn_stem, 1..3 = lmfa{2,3,4} + rrcv{2,3,4}
        self.part_idx = [[i for i, n in enumerate(names) if n.startswith("cnn_stem.")]]
        for k in (2, 3, 4):
            self.part_idx.append([i for i, n in enumerate(names) if n.startswith((f"lmfa{k}.", f"rrcv{k}."))])

    def part_tensors(self, part):
        return [self.tensors[i] for i in self.part_idx[part]]


class LateralPartFn(torch.autograd.Function):
    """One phase of the lateral path.  part 0: image -> token (the stem's feature maps stay in state.saved; the token only carries
    the autograd dependency); part k = 1..3: token -> R_{k+1} (LMFAdapter + RRCV of stage k + 1).  Separate autograd nodes let the
    adapters' backward start as soon as their dR is known and the stem's backward run last, all on the lateral side stream, next to
    the token path's blocks (H:1236-1247 computes the same three tensors up front)."""

    @staticmethod
    def forward(ctx, inp, state: LateralState, part: int, *tensors):
        cfg = state.cfg
        dev = state.img.device
        scratch = _scratch_buf(dev, state.scratch_bytes)
        out = None
        Rp = [None, None, None]
        if part == 0:
            out = torch.empty(1, dtype=torch.float32, device=dev)    # never read: it only orders the stem's backward after the adapters'
            ctx.set_materialize_grads(False)                          # ... whose token gradient is None (no ATen fill / add kernels)
        else:
            N = cfg.grid * cfg.grid
            out = torch.empty(cfg.batch, N, cfg.dim, dtype=torch.float32, device=dev)
            Rp[part - 1] = out.data_ptr()
        check(lib.qavit_lateral_forward_parts(C.byref(cfg), state.params, state.img.data_ptr(), Rp[0], Rp[1], Rp[2],
                                              state.saved.data_ptr(), scratch.data_ptr(), _stream(), 1 << part))
        ctx.state, ctx.part, ctx.tensors = state, part, tensors
        return out

    @staticmethod
    def backward(ctx, dout):
        state, part = ctx.state, ctx.part
        cfg, dev = state.cfg, state.img.device
        idx = state.part_idx[part]
        skip = {j for j, i in enumerate(idx) if i in state.buffers}
        arr, views, direct = _grad_arrays(ctx.tensors, skip, dev)
        full = (C.c_void_p * len(state.tensors))()
        for j, i in enumerate(idx):
            full[i] = arr[j]
        scratch = _scratch_buf(dev, state.scratch_bytes)
        dR = [None, None, None]
        parts = 1 << part
        if part == 0:
            for k in range(3):                      # an adapter that took no part in the loss: its feature-map gradient is zero
                if not state.bwd_done[k]:
                    check(lib.qavit_lateral_backward_parts(C.byref(cfg), state.params, full, state.img.data_ptr(), None, None, None,
                                                           state.saved.data_ptr(), scratch.data_ptr(), _stream(), 2 << k))
        else:
            dR[part - 1] = dout.float().contiguous()
            state.bwd_done[part - 1] = True
        check(lib.qavit_lateral_backward_parts(C.byref(cfg), state.params, full, state.img.data_ptr(), _ptr(dR[0]), _ptr(dR[1]),
                                               _ptr(dR[2]), state.saved.data_ptr(), scratch.data_ptr(), _stream(), parts))
        if part == 0:
            state.saved = None
        _notify(direct)
        return (None, None, None, *views)


class SplitFusionFn(torch.autograd.Function):
    """SplitFusion.forward -- HQAViT_CIFAR100.py:945-963."""

    @staticmethod
    def forward(ctx, T_in, R, cfg: SplitFusionCfg, *tensors):
        _require_cuda(T_in, "SplitFusion input")
        T_in = T_in.detach().float().contiguous()
        R = R.detach().float().contiguous()
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_splitfusion_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        saved = torch.empty(saved_b.value, dtype=torch.uint8, device=T_in.device)
        scratch = _scratch_buf(T_in.device, scratch_b.value)
        params = _param_array(tensors, len(tensors))
        out = torch.empty_like(T_in)
        rng = rng_state(T_in.device) if (cfg.train and cfg.drop_p > 0) else None
        check(lib.qavit_splitfusion_forward(C.byref(cfg), params, _ptr(rng), T_in.data_ptr(), R.data_ptr(), out.data_ptr(),
                                            saved.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.cfg, ctx.saved_buf, ctx.T_in, ctx.R, ctx.params_arr, ctx.tensors = cfg, saved, T_in, R, params, tensors
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg, T_in, R = ctx.cfg, ctx.T_in, ctx.R
        dout = dout.float().contiguous()
        grads_arr, views, direct = _grad_arrays(ctx.tensors, (), T_in.device)
        saved_b, scratch_b = C.c_size_t(0), C.c_size_t(0)
        check(lib.qavit_splitfusion_workspace(C.byref(cfg), C.byref(saved_b), C.byref(scratch_b)))
        scratch = _scratch_buf(T_in.device, scratch_b.value)
        dT = torch.empty_like(T_in)
        dR = torch.empty_like(R)
        check(lib.qavit_splitfusion_backward(C.byref(cfg), ctx.params_arr, grads_arr, T_in.data_ptr(), R.data_ptr(), dout.data_ptr(),
                                             dT.data_ptr(), dR.data_ptr(), ctx.saved_buf.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.saved_buf = None
        _notify(direct)
        return (dT, dR, None, *views)


_ce_err = {}


def _ce_err_flag(device) -> torch.Tensor:
    """Device int raised by the loss kernel when a label is out of range (read by check_labels())."""
    device = torch.device(device)
    t = _ce_err.get(device)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=device)
        _ce_err[device] = t
    return t


def check_labels(device=None) -> None:
    """Raise (like torch's device-side assert, but recoverable) if any cross_entropy() call since the last check saw a
    target outside [0, classes).  One host sync; the train step itself never synchronises."""
    for dev, t in list(_ce_err.items()):
        if device is not None and torch.device(device) != dev:
            continue
        if int(t.item()) != 0:
            t.zero_()
            raise RuntimeError("qavit_b200.cross_entropy: target out of range [0, num_classes)")


class CrossEntropyFn(torch.autograd.Function):
    """CrossEntropyLoss(label_smoothing) and its two-target mixup form -- HQAViT_CIFAR100.py:1373, 1404-1408."""

    @staticmethod
    def forward(ctx, logits, ya, yb, lam, smoothing):
        _require_cuda(logits, "logits")
        logits = logits.detach().float().contiguous()
        B, ncls = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        row_loss = torch.empty(max(B, 1), dtype=torch.float32, device=logits.device)
        lam_dev = lam if isinstance(lam, torch.Tensor) else None      # device scalar: graph-replayable mixup weight
        if lam_dev is not None and (not lam_dev.is_cuda or lam_dev.dtype != torch.float32):
            raise RuntimeError("qavit_b200.cross_entropy: a tensor `lam` must be a CUDA float32 scalar")
        check(lib.qavit_cross_entropy(logits.data_ptr(), ya.data_ptr(), _ptr(yb), 1.0 if lam_dev is not None else float(lam),
                                      _ptr(lam_dev), B, ncls, float(smoothing), loss.data_ptr(), dlogits.data_ptr(),
                                      row_loss.data_ptr(), _ce_err_flag(logits.device).data_ptr(), _stream()))
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        dloss = dloss.float().contiguous()
        out = torch.empty_like(dlogits)
        check(lib.qavit_scale_by_scalar(dlogits.data_ptr(), dloss.data_ptr(), dlogits.numel(), out.data_ptr(), _stream()))
        return out, None, None, None, None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, label_smoothing: float = 0.0,
                  target_b: Optional[torch.Tensor] = None, lam=1.0) -> torch.Tensor:
    """Drop-in for ``nn.CrossEntropyLoss(label_smoothing=...)(logits, target)``; with ``target_b`` the
    ``lam * CE(a) + (1 - lam) * CE(b)`` mixup form of the reference's train loop.  ``lam`` may be a CUDA float32 scalar
    tensor (read on the device: a captured step can change it between replays).  Out-of-range targets raise a device
    flag that ``check_labels()`` turns into a RuntimeError; the loss is summed in a fixed order (bitwise reproducible)."""
    ya = target.to(torch.int64).contiguous()
    yb = None if target_b is None else target_b.to(torch.int64).contiguous()
    return CrossEntropyFn.apply(logits, ya, yb, lam, label_smoothing)


# ------------------------------------------------------------------------------------------------ batch augmentation (8(f)-3)
def rand_bbox(size, lam, rng=None):
    """``rand_bbox`` of the reference (HQAViT_CIFAR100.py:1343-1360): centre uniform, side = sqrt(1 - lam) of the image."""
    import numpy as np
    rng = np.random if rng is None else rng
    W, H = size[3], size[2]
    cut_rat = np.sqrt(1.0 - lam)
    cut_w, cut_h = int(W * cut_rat), int(H * cut_rat)
    cx, cy = rng.randint(W), rng.randint(H)
    return (int(np.clip(cx - cut_w // 2, 0, W)), int(np.clip(cy - cut_h // 2, 0, H)),
            int(np.clip(cx + cut_w // 2, 0, W)), int(np.clip(cy + cut_h // 2, 0, H)))


def batch_mix(inputs: torch.Tensor, perm: torch.Tensor, mode: str, lam: float = 1.0, box=(0, 0, 0, 0)) -> torch.Tensor:
    """One kernel for ``inputs[:, :, y1:y2, x1:x2] = inputs[perm, :, y1:y2, x1:x2]`` (mode 'cutmix') or
    ``lam * inputs + (1 - lam) * inputs[perm]`` (mode 'mixup') -- HQAViT_CIFAR100.py:1383-1397."""
    _require_cuda(inputs, "image batch")
    x = inputs.float().contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty_like(x)
    x1, y1, x2, y2 = box
    check(lib.qavit_batch_mix(x.data_ptr(), out.data_ptr(), perm.to(torch.int64).contiguous().data_ptr(), B, Cc, H, W,
                              1 if mode == "cutmix" else 2, int(x1), int(y1), int(x2), int(y2), float(lam), 1.0 - float(lam), _stream()))
    return out


def mix_batch(inputs: torch.Tensor, targets: torch.Tensor, config, rng=None):
    """The augmentation branch of ``train_epoch`` (HQAViT_CIFAR100.py:1379-1399) with the pixel work on the device:
    returns ``(inputs, targets_a, targets_b, lam, use_mix)``; the loss is then
    ``cross_entropy(logits, targets_a, label_smoothing, target_b=targets_b, lam=lam)`` (H:1404-1408).  Host-side random
    draws (np.random.rand / beta / randperm) follow the reference's order."""
    import numpy as np
    rng = np.random if rng is None else rng
    if getattr(config, "use_cutmix", False) and rng.rand() < config.mix_prob:
        perm = torch.randperm(inputs.size(0)).to(inputs.device)
        x1, y1, x2, y2 = rand_bbox(inputs.size(), rng.beta(config.cutmix_alpha, config.cutmix_alpha), rng)
        out = batch_mix(inputs, perm, "cutmix", box=(x1, y1, x2, y2))
        lam = 1.0 - ((x2 - x1) * (y2 - y1) / float(inputs.size(3) * inputs.size(2)))
        return out, targets, targets[perm], lam, "cutmix"
    if getattr(config, "use_mixup", False) and rng.rand() < config.mix_prob:
        perm = torch.randperm(inputs.size(0)).to(inputs.device)
        lam = float(rng.beta(config.mixup_alpha, config.mixup_alpha))
        return batch_mix(inputs, perm, "mixup", lam=lam), targets, targets[perm], lam, "mixup"
    return inputs, targets, targets, 1.0, None
