"""The fused TokenLearner / TokenUpMix kernels of bf16 runs (tokens_fused.cu) against the reference expressions
(HQAViT_CIFAR100.py:985-1002 and 1016-1031) evaluated by torch autograd in fp64, through the C ABI.  The kernels use bf16
operand PAIRS (hi + lo) with fp32 accumulation, so they are held to 2e-4 -- fifty times tighter than the north star's bf16
tolerance: the block wrapper has no residual path around it, its rounding lands directly on the fp32 token stream."""
import ctypes as C

import pytest
import torch

from util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-4


def _call(op, B, N, Cc, ins, outs):
    from qavit_b200 import _lib as L
    a = (C.c_void_p * len(ins))(*[t.data_ptr() for t in ins])
    o = (C.c_void_p * len(outs))(*[t.data_ptr() for t in outs])
    L.check(L.lib.qavit_test_tokens_fused(op, B, N, Cc, a, o, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


def _ref_token_learner(x, lw, lb, W, b):
    ln = torch.nn.functional.layer_norm(x, (x.shape[-1],), lw, lb, 1e-5)
    logits = ln @ W.t() + b                         # [B, N, 16]
    S = torch.softmax(logits, dim=1)                # over the token axis, H:996
    return S, torch.bmm(S.transpose(1, 2), x)       # [B, 16, C], H:1000


def _ref_upmix(xc, W, b, lw, lb):
    up = (xc.transpose(1, 2) @ W.t() + b).transpose(1, 2)     # Linear over the token axis, H:1024-1026
    return torch.nn.functional.layer_norm(up, (up.shape[-1],), lw, lb, 1e-5), up


@pytest.mark.parametrize("B,N", [(5, 64), (300, 64), (3, 16), (7, 48)])
def test_token_learner_fused_forward_and_backward(B, N):
    Cc = 192
    g = torch.Generator().manual_seed(B * 100 + N)
    x = (torch.randn(B, N, Cc, generator=g) * 1.3 + 0.2 * torch.randn(B, N, 1, generator=g)).cuda()
    lw = (1 + 0.2 * torch.randn(Cc, generator=g)).cuda()
    lb = (0.1 * torch.randn(Cc, generator=g)).cuda()
    W = (torch.randn(16, Cc, generator=g) * 0.15).cuda()
    b = (0.1 * torch.randn(16, generator=g)).cuda()
    dxc = torch.randn(B, 16, Cc, generator=g).cuda()
    S = torch.empty(B, N, 16, device="cuda")
    xc = torch.empty(B, 16, Cc, device="cuda")
    Z = torch.empty(B, N, 16, device="cuda")
    _call(0, B, N, Cc, [x, lw, lb, W, b], [S, xc, Z])
    leaves = [t.double().requires_grad_(True) for t in (x, lw, lb, W, b)]
    S_ref, xc_ref = _ref_token_learner(*leaves)
    assert rel_l2(S, S_ref) < TOL, rel_l2(S, S_ref)
    assert rel_l2(xc, xc_ref) < TOL, rel_l2(xc, xc_ref)
    gx, glw, glb, gW, gb = torch.autograd.grad((xc_ref * dxc.double()).sum(), leaves)
    dx = torch.empty_like(x)
    dW, db, dlw, dlb = torch.zeros_like(W), torch.zeros_like(b), torch.zeros_like(lw), torch.zeros_like(lb)
    _call(1, B, N, Cc, [x, S, dxc, lw, lb, W, Z], [dx, dW, db, dlw, dlb])
    assert rel_l2(dx, gx) < TOL, rel_l2(dx, gx)
    assert rel_l2(dW, gW) < TOL, rel_l2(dW, gW)
    assert rel_l2(dlw, glw) < TOL, rel_l2(dlw, glw)
    # the gate bias AND the LayerNorm bias only add a per-slot constant to the logits, which the softmax over tokens removes: both
    # gradients are exactly zero in exact arithmetic (fp64: ~1e-15) and are bounded absolutely here
    assert gb.abs().max().item() < 1e-9 and glb.abs().max().item() < 1e-9
    assert db.abs().max().item() < 1e-4 * gW.abs().max().item() * Cc ** 0.5
    assert dlb.abs().max().item() < 1e-4 * glw.abs().max().item()


@pytest.mark.parametrize("B,N", [(5, 64), (300, 64), (3, 16), (7, 32)])
def test_token_upmix_fused_forward_and_backward(B, N):
    Cc = 192
    g = torch.Generator().manual_seed(B * 100 + N + 1)
    xc = (torch.randn(B, 16, Cc, generator=g) * 1.1).cuda()
    W = (torch.randn(N, 16, generator=g) * 0.3).cuda()
    b = (0.2 * torch.randn(N, generator=g)).cuda()
    lw = (1 + 0.2 * torch.randn(Cc, generator=g)).cuda()
    lb = (0.1 * torch.randn(Cc, generator=g)).cuda()
    dout = torch.randn(B, N, Cc, generator=g).cuda()
    out = torch.empty(B, N, Cc, device="cuda")
    stats = torch.empty(B * N, 2, device="cuda")
    _call(2, B, N, Cc, [xc, W, b, lw, lb], [out, stats])
    leaves = [t.double().requires_grad_(True) for t in (xc, W, b, lw, lb)]
    out_ref, up_ref = _ref_upmix(*leaves)
    assert rel_l2(out, out_ref) < TOL, rel_l2(out, out_ref)
    mean_ref = up_ref.mean(-1).reshape(-1)
    rstd_ref = (up_ref.var(-1, unbiased=False) + 1e-5).rsqrt().reshape(-1)
    assert rel_l2(stats[:, 0], mean_ref, floor=1e-3) < 1e-4 and rel_l2(stats[:, 1], rstd_ref) < 1e-4
    gxc, gW, gb, glw, glb = torch.autograd.grad((out_ref * dout.double()).sum(), leaves)
    dxc = torch.empty_like(xc)
    dW, dlw, dlb = torch.zeros_like(W), torch.zeros_like(lw), torch.zeros_like(lb)
    _call(3, B, N, Cc, [xc, dout, stats, W, b, lw], [dxc, dW, dlw, dlb])
    assert rel_l2(dxc, gxc) < TOL, rel_l2(dxc, gxc)
    assert rel_l2(dW, gW) < TOL, rel_l2(dW, gW)
    assert rel_l2(dlw, glw) < TOL and rel_l2(dlb, glb) < TOL, (rel_l2(dlw, glw), rel_l2(dlb, glb))
    assert gb.abs().max().item() < 1e-9 * max(1.0, gW.abs().max().item())      # exactly zero in exact arithmetic: we add nothing


@pytest.mark.parametrize("R", [64, 1000, 20000])
def test_branch_norm_compress_fused_forward_and_backward(R):
    """cmp_fused.cu: fused[:, 48 i : 48 i + 48] = alpha_i (LN_i(x_i) W_i^T + b_i) for the four branches (H:1074-1079) and its
    backward, against torch autograd in fp64 on the same bf16 inputs.  Outputs are bf16: tolerance = bf16 rounding of the result."""
    from qavit_b200 import _lib as L
    Cc, Cd = 192, 48
    g = torch.Generator().manual_seed(R)
    xs = [(torch.randn(R, Cc, generator=g) * (1 + 0.5 * i) + 0.3).bfloat16().cuda() for i in range(4)]
    gam = [(1 + 0.2 * torch.randn(Cc, generator=g)).cuda() for _ in range(4)]
    bet = [(0.1 * torch.randn(Cc, generator=g)).cuda() for _ in range(4)]
    Ws = [(torch.randn(Cd, Cc, generator=g) * 0.1).cuda() for _ in range(4)]
    bs = [(0.1 * torch.randn(Cd, generator=g)).cuda() for _ in range(4)]
    alpha = torch.softmax(torch.randn(4, generator=g), 0).cuda()
    dfused = torch.randn(R, Cc, generator=g).bfloat16().cuda()
    fused = torch.empty(R, Cc, dtype=torch.bfloat16, device="cuda")
    stats = [torch.empty(R, 2, device="cuda") for _ in range(4)]
    ins = xs + gam + bet + Ws + bs + [alpha]
    a = (C.c_void_p * len(ins))(*[t.data_ptr() for t in ins])
    o = (C.c_void_p * 5)(*[t.data_ptr() for t in [fused] + stats])
    L.check(L.lib.qavit_test_cmp_fused(0, R, a, o, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    leaves = {k: [t.double().requires_grad_(True) for t in v] for k, v in dict(x=xs, g=gam, b=bet, W=Ws, c=bs).items()}
    ref = torch.cat([alpha[i].double() * (torch.nn.functional.layer_norm(leaves["x"][i], (Cc,), leaves["g"][i], leaves["b"][i], 1e-5)
                                         @ leaves["W"][i].t() + leaves["c"][i]) for i in range(4)], dim=1)
    assert rel_l2(fused.float(), ref) < 4e-3, rel_l2(fused.float(), ref)
    for i in range(4):
        assert rel_l2(stats[i][:, 0], xs[i].double().mean(-1), floor=1e-3) < 1e-5
        assert rel_l2(stats[i][:, 1], (xs[i].double().var(-1, unbiased=False) + 1e-5).rsqrt()) < 1e-5
    flat = leaves["x"] + leaves["g"] + leaves["b"] + leaves["W"] + leaves["c"]
    grads = torch.autograd.grad((ref * dfused.double()).sum(), flat)
    gx, gg, gb, gW, gc = (grads[4 * k:4 * k + 4] for k in range(5))
    dx = [torch.empty(R, Cc, dtype=torch.bfloat16, device="cuda") for _ in range(4)]
    dW = [torch.zeros(Cd, Cc, device="cuda") for _ in range(4)]
    db = [torch.zeros(Cd, device="cuda") for _ in range(4)]
    dg = [torch.zeros(Cc, device="cuda") for _ in range(4)]
    dbe = [torch.zeros(Cc, device="cuda") for _ in range(4)]
    ins = xs + stats + gam + bet + Ws + [alpha, dfused]
    a = (C.c_void_p * len(ins))(*[t.data_ptr() for t in ins])
    outs = dx + dW + db + dg + dbe
    o = (C.c_void_p * len(outs))(*[t.data_ptr() for t in outs])
    L.check(L.lib.qavit_test_cmp_fused(1, R, a, o, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for i in range(4):
        assert rel_l2(dx[i].float(), gx[i]) < 6e-3, (i, rel_l2(dx[i].float(), gx[i]))         # bf16 output, bf16 alpha W
        assert rel_l2(dW[i], gW[i]) < 4e-3, (i, rel_l2(dW[i], gW[i]))
        assert rel_l2(db[i], gc[i]) < 4e-3, (i, rel_l2(db[i], gc[i]))      # column sums ride the MMA through a (1 / rstd) hi + lo column pair
        assert rel_l2(dg[i], gg[i]) < 4e-3 and rel_l2(dbe[i], gb[i]) < 4e-3, (i, rel_l2(dg[i], gg[i]), rel_l2(dbe[i], gb[i]))
