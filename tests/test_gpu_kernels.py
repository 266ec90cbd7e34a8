"""GPU unit tests of the GEMM flavours and the small standalone ops, through the C ABI (ctypes)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from qavit_b200 import _lib
    return _lib


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,N,K", [(300, 192, 192), (64, 576, 192), (1000, 48, 192), (130, 100, 192), (77, 192, 96)])
def test_simt_gemm_nt_and_tn(M, N, K):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    Cc = torch.empty(M, N, device="cuda")
    L.check(L.lib.qavit_test_gemm_nt(0, A.data_ptr(), K, M, N, K, W.data_ptr(), None, b.data_ptr(), Cc.data_ptr(), 1, _s()))
    ref = (A.double() @ W.double().t() + b.double()).float()
    assert (Cc - ref).abs().max().item() < 1e-4
    dY = torch.randn(M, N, device="cuda", generator=g)
    dW = torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    L.check(L.lib.qavit_test_gemm_tn(0, dY.data_ptr(), N, A.data_ptr(), K, M, N, K, dW.data_ptr(), db.data_ptr(), _s()))
    refW = (dY.double().t() @ A.double()).float()
    assert (dW - refW).abs().max().item() < 1e-3 * refW.abs().max().item()
    assert (db - dY.sum(0)).abs().max().item() < 1e-3 * dY.sum(0).abs().max().item()


TC_NT = [(128, 192, 192), (4096, 192, 192), (1000, 576, 192), (333, 96, 192), (256, 192, 96), (64, 48, 192),
         (512, 16, 192), (2048, 208, 192), (640, 192, 576), (300, 192, 48), (1024, 192, 16), (160, 384, 192)]


@pytest.mark.parametrize("M,N,K", TC_NT)
def test_tcgen05_gemm_nt(M, N, K):
    """bf16 operands, fp32 accumulation: equal to an fp64 product of the same bf16-rounded operands to ~1e-5."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    Wb = W.bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    Cc = torch.full((M, N), float("nan"), device="cuda")
    L.check(L.lib.qavit_test_gemm_nt(1, A.data_ptr(), K, M, N, K, W.data_ptr(), Wb.data_ptr(), b.data_ptr(), Cc.data_ptr(), 1, _s()))
    torch.cuda.synchronize()
    ref = (A.double() @ Wb.double().t() + b.double()).float()
    err = (Cc - ref).abs().max().item()
    assert err < 2e-4, f"max abs err {err} (ref max {ref.abs().max().item()})"


TC_TN = [(700, 64, 288), (1000, 256, 1024), (900, 192, 768), (512, 192, 384), (4096, 192, 192), (1000, 576, 192), (333, 96, 192), (2048, 192, 96), (640, 48, 192), (512, 16, 192),
         (64, 192, 192), (3000, 384, 192), (130, 192, 48)]


@pytest.mark.parametrize("M,N,K", TC_TN)
def test_tcgen05_gemm_tn(M, N, K):
    """dW[N, K] += dY[M, N]^T X[M, K] with MN-major UMMA operands (no transposed copies)."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(2)
    dY = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    X = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    dW = torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    L.check(L.lib.qavit_test_gemm_tn(1, dY.data_ptr(), N, X.data_ptr(), K, M, N, K, dW.data_ptr(), db.data_ptr(), _s()))
    torch.cuda.synchronize()
    ref = (dY.double().t() @ X.double()).float()
    err = (dW - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err} (ref max {ref.abs().max().item()})"
    assert (db - dY.float().sum(0)).abs().max().item() < 1e-2


@pytest.mark.parametrize("M,N,K", [(300, 192, 192), (1000, 1024, 256), (640, 256, 1024), (130, 64, 288)])
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_tcgen05_gemm_fused_epilogues(M, N, K, mode):
    """GELU dual output, bf16 residual and GELU-backward multiply epilogues of the tcgen05 NT GEMM (lateral path)."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    Wb = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    aux = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    C2 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.qavit_test_gemm_epi(A.data_ptr(), K, M, N, K, Wb.data_ptr(), bias.data_ptr(), C.data_ptr(), C2.data_ptr(), mode,
                                      aux.data_ptr(), _s()))
    torch.cuda.synchronize()
    pre = A.double() @ Wb.double().t() + bias.double()
    tol = 1.2e-2   # bf16 output rounding (2^-8 relative) on values of magnitude <= ~4
    if mode == 1:
        assert (C.double() - pre).abs().max().item() < tol * 4
        assert (C2.double() - torch.nn.functional.gelu(pre)).abs().max().item() < tol * 4
    elif mode == 2:
        assert (C2.double() - (pre + aux.double())).abs().max().item() < tol * 6
    elif mode == 4:     # forward stores gelu'(pre) (what backward multiplies by) next to gelu(pre)
        dg = 0.5 * (1 + torch.erf(pre / 2 ** 0.5)) + pre * torch.exp(-0.5 * pre * pre) / (2 * torch.pi) ** 0.5
        assert (C.double() - dg).abs().max().item() < tol
        assert (C2.double() - torch.nn.functional.gelu(pre)).abs().max().item() < tol * 4
    elif mode == 5:
        assert (C.double() - pre * aux.double()).abs().max().item() < tol * 6
    else:
        x = aux.double()
        dg = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
        assert (C.double() - pre * dg).abs().max().item() < tol * 6


def test_tcgen05_gemm_nt_strided_slice():
    """A operand that is a 48-column slice of a 192-wide buffer (the compress dX GEMM): TMA must zero-fill beyond it."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 500, 192, 48
    buf = torch.randn(M, 192, device="cuda", generator=g).bfloat16()
    W = torch.randn(N, K, device="cuda", generator=g) / 7
    Wb = W.bfloat16()
    Cc = torch.empty(M, N, device="cuda")
    A = buf[:, 48:96]
    L.check(L.lib.qavit_test_gemm_nt(1, A.data_ptr(), 192, M, N, K, W.data_ptr(), Wb.data_ptr(), None, Cc.data_ptr(), 1, _s()))
    ref = (A.double() @ Wb.double().t()).float()
    assert (Cc - ref).abs().max().item() < 2e-4


def test_cross_entropy_matches_torch():
    import qavit_b200 as Q
    g = torch.Generator(device="cuda").manual_seed(4)
    logits = torch.randn(37, 100, device="cuda", generator=g, requires_grad=True)
    ya = torch.randint(0, 100, (37,), device="cuda", generator=g)
    yb = torch.randint(0, 100, (37,), device="cuda", generator=g)
    loss = Q.cross_entropy(logits, ya, label_smoothing=0.12)
    loss.backward()
    l2 = logits.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(l2, ya, label_smoothing=0.12)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert (logits.grad - l2.grad).abs().max().item() < 1e-6
    # mixup / cutmix two-target form, H:1407
    loss2 = Q.cross_entropy(logits.detach(), ya, label_smoothing=0.12, target_b=yb, lam=0.3)
    ce = torch.nn.CrossEntropyLoss(label_smoothing=0.12)
    ref2 = 0.3 * ce(l2.detach(), ya) + 0.7 * ce(l2.detach(), yb)
    assert abs(loss2.item() - ref2.item()) < 1e-5


def test_clip_and_adamw_match_torch():
    """H:1413-1439: per-parameter clip (dwconv / cnn_stem names), global clip, AdamW; params without grad are skipped."""
    import qavit_b200 as Q
    torch.manual_seed(5)
    shapes = {"a.weight": (33, 7), "cnn_stem.x.weight": (64, 3), "b.dwconv.weight": (96, 9), "c.bias": (5,), "never.used": (11,)}
    ours = {n: torch.nn.Parameter(torch.randn(s, device="cuda")) for n, s in shapes.items()}
    ref = {n: torch.nn.Parameter(p.detach().clone()) for n, p in ours.items()}
    opt = Q.FusedAdamW(list(ours.items()), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06, max_grad_norm=0.5)
    ropt = torch.optim.AdamW(ref.values(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06)
    for step in range(3):
        opt.zero_grad()
        ropt.zero_grad(set_to_none=True)
        gs = {n: torch.randn(s, device="cuda") * (3.0 if "dwconv" in n else 0.2) for n, s in shapes.items() if n != "never.used"}
        for n, gval in gs.items():
            ours[n].grad.copy_(gval)
            ref[n].grad = gval.clone()
        opt.set_grad_mask([n != "never.used" for n in shapes])
        for n, p in ref.items():
            if p.grad is not None and ("cnn_stem" in n or "dwconv" in n):
                torch.nn.utils.clip_grad_norm_([p], 0.1)
        rnorm = torch.nn.utils.clip_grad_norm_(ref.values(), 0.5)
        norm = opt.clip()
        assert abs(norm.item() - rnorm.item()) < 1e-5 * max(1.0, rnorm.item())
        for n in gs:
            assert (ours[n].grad - ref[n].grad).abs().max().item() < 1e-6, n
        for g_ in opt.param_groups:
            g_["lr"] = 6e-4 * (step + 1)
            g_["betas"] = (0.95 - 0.01 * step, 0.999)
        for g_ in ropt.param_groups:
            g_["lr"] = 6e-4 * (step + 1)
            g_["betas"] = (0.95 - 0.01 * step, 0.999)
        opt.step()
        ropt.step()
        for n in shapes:
            assert (ours[n].data - ref[n].data).abs().max().item() < 2e-6, (n, step)


@pytest.mark.parametrize("K", [3, 5, 7])
@pytest.mark.parametrize("C,B", [(64, 37), (128, 16), (256, 5)])
def test_depthwise_8x8_tensor_core_path_matches_conv2d(K, C, B):
    """bf16 depthwise k x k convolution on 8 x 8 channels-last maps (dwconv_mma.cu: the stencil as a per-channel 64 x 64 Toeplitz
    product on mma.sync) and its input gradient (flipped taps) against torch.nn.functional.conv2d in fp64 on the bf16-rounded
    operands; B = 37 / 5 leave a partial 16-image tile."""
    import torch.nn.functional as F
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(K * 100 + C)
    x = torch.randn(B, 8, 8, C, device="cuda", generator=g).bfloat16()               # channels-last rows [B * 64, C]
    w = torch.randn(C, K, K, device="cuda", generator=g) / K
    bias = torch.randn(C, device="cuda", generator=g)
    y = torch.zeros(B, 8, 8, C, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.qavit_dwconv_forward(x.data_ptr(), 1, B, 8, 8, C, K, w.data_ptr(), bias.data_ptr(), y.data_ptr(), _s()))
    wq = w.bfloat16().double()                                                       # the kernel rounds its taps to bf16
    xr = x.double().permute(0, 3, 1, 2)
    ref = F.conv2d(xr, wq.view(C, 1, K, K), bias.double(), padding=K // 2, groups=C).permute(0, 2, 3, 1)
    err = (y.double() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err                        # bf16 output rounding
    # backward: dx = conv(dy, flipped taps); dw / dbias from the CUDA-core kernel
    dy = torch.randn(B, 8, 8, C, device="cuda", generator=g).bfloat16()
    dx = torch.zeros_like(dy)
    dw = torch.zeros(C, K, K, device="cuda")
    db = torch.zeros(C, device="cuda")
    L.check(L.lib.qavit_dwconv_backward(x.data_ptr(), dy.data_ptr(), 1, B, 8, 8, C, K, w.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                        db.data_ptr(), _s()))
    torch.cuda.synchronize()
    xg = xr.clone().requires_grad_(True)
    wg = wq.clone().view(C, 1, K, K).requires_grad_(True)
    out = F.conv2d(xg, wg, None, padding=K // 2, groups=C)
    out.backward(dy.double().permute(0, 3, 1, 2))
    rdx = xg.grad.permute(0, 2, 3, 1)
    assert (dx.double() - rdx).abs().max().item() < 2e-2 * max(1.0, rdx.abs().max().item())
    assert (dw.double() - wg.grad.view(C, K, K)).abs().max().item() < 2e-2 * max(1.0, wg.grad.abs().max().item())


@pytest.mark.parametrize("side,C,B,bias", [(4, 96, 37, False), (4, 96, 300, True), (4, 32, 9, True), (4, 64, 5, False), (4, 128, 7, True),
                                           (8, 96, 37, False), (8, 96, 700, True), (8, 48, 3, True), (8, 128, 5, False)])
def test_ffn_mid_fused_matches_autograd(side, C, B, bias):
    """CCF-FFN mid-section as one kernel per direction (ffn_mid.cu; H:704-709): GELU -> LayerNorm -> depthwise 3x3 (+ bias) * scale
    -> LayerNorm on 4 x 4 token maps (register kernel) and 8 x 8 maps (shared-memory tile kernel), against torch autograd in fp64 on
    the bf16-rounded inputs.  Outputs are bf16 (2^-9 relative rounding); parameter gradients are fp32 sums and are held to 1e-3.
    The small B leave a partial group of images / idle CTAs; B = 700 makes a CTA walk several images."""
    import ctypes as Ct
    import torch.nn.functional as F
    from util import rel_l2
    L = _lib()
    T = side * side
    g = torch.Generator(device="cuda").manual_seed(C * 7 + B + side)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    h_pre = (rn(B, T, C) * 1.2 + 0.1).bfloat16()
    g1, b1, g2, b2 = 1 + 0.2 * rn(C), 0.1 * rn(C), 1 + 0.2 * rn(C), 0.1 * rn(C)
    w, cb, sc = rn(C, 3, 3) / 3, 0.2 * rn(C), 1 + 0.3 * rn(C)
    d_hn2 = rn(B, T, C).bfloat16()
    hn2 = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
    st1, st2 = torch.empty(B * T, 2, device="cuda"), torch.empty(B * T, 2, device="cuda")

    def call(op, ins, outs):
        a = (Ct.c_void_p * len(ins))(*[t.data_ptr() if t is not None else None for t in ins])
        o = (Ct.c_void_p * len(outs))(*[t.data_ptr() if t is not None else None for t in outs])
        L.check(L.lib.qavit_test_ffn_mid(op, B, side, C, a, o, _s()))
        torch.cuda.synchronize()

    params = [g1, b1, w, cb if bias else None, sc, g2, b2]
    call(0, [h_pre] + params, [hn2, st1, st2])
    leaves = [t.double().requires_grad_(True) for t in (h_pre, g1, b1, w, cb, sc, g2, b2)]
    x, G1, B1, W, CB, SC, G2, B2 = leaves
    h = F.layer_norm(F.gelu(x), (C,), G1, B1, 1e-5)
    img = h.transpose(1, 2).reshape(B, C, side, side)
    img = F.conv2d(img, W.view(C, 1, 3, 3), CB if bias else None, padding=1, groups=C) * SC.view(1, C, 1, 1)
    ref = F.layer_norm(img.flatten(2).transpose(1, 2), (C,), G2, B2, 1e-5)
    assert rel_l2(hn2, ref) < 4e-3, rel_l2(hn2, ref)                                  # bf16 output
    grads = torch.autograd.grad((ref * d_hn2.double()).sum(), leaves, allow_unused=True)
    d_hpre = torch.empty_like(h_pre)
    dg1, db1, dg2, db2, dsc, dcb = (torch.zeros(C, device="cuda") for _ in range(6))
    dw = torch.zeros(C, 3, 3, device="cuda")
    call(1, [h_pre] + params + [d_hn2, st1, st2], [d_hpre, dg1, db1, dw, dcb if bias else None, dsc, dg2, db2])
    assert rel_l2(d_hpre, grads[0]) < 4e-3, rel_l2(d_hpre, grads[0])
    for name, ours, r in (("dg1", dg1, grads[1]), ("db1", db1, grads[2]), ("dw", dw, grads[3]), ("dscale", dsc, grads[5]),
                          ("dg2", dg2, grads[6]), ("db2", db2, grads[7])) + ((("dbias", dcb, grads[4]),) if bias else ()):
        assert rel_l2(ours, r) < 1e-3, (name, rel_l2(ours, r))
