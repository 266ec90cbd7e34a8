"""Train-mode dropout / DropPath of the quad block (H:256-264, 416-465, 648-656, 697-710, 1066-1083) through the C ABI.

The kernels draw their masks from a counter-based generator (Philox4x32-7 keyed by {seed, offset, site, element id}), so
the masks are a pure function that `tests/dropout_masks.py` restates on the host; the oracle takes those masks as explicit
keep-scale tensors.  That turns the dropout run into an element-wise parity test: fp32 run 1e-4 on the block output and
gradients (same gates as the dropout-free block test), bf16 run at the bf16 gates -- plus rate / unbiasedness / replay
checks of the generator itself."""
import numpy as np
import pytest
import torch

from oracle import qavit_oracle as O
from dropout_masks import block_masks
from util import build_model, rel_l2, rel_max

pytestmark = pytest.mark.gpu

SEED = 0x1234ABCD5678


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _set_rng(seed, offset=0):
    import qavit_b200.functional as QF
    st = QF.rng_state(torch.device("cuda", torch.cuda.current_device()))
    st.copy_(torch.tensor([seed, offset], dtype=torch.int64))
    return st


def _oracle_block(ocfg, sd, prefix, wrapped, x, wgt, masks):
    keys = [k for k in O.trainable_keys(ocfg) if k.startswith(prefix + ".") or k.startswith("global_bank.")]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    full = dict(sd)
    full.update(leaves)
    xl = x.detach().clone().requires_grad_(True)
    bank = O.Bank(full, ocfg)
    out = (O.wrapped_block if wrapped else O.quad_block)(xl, full, prefix, ocfg, bank, True, masks)
    gs = torch.autograd.grad((out * wgt).sum(), [xl] + list(leaves.values()), allow_unused=True)
    return out.detach(), gs[0], dict(zip(keys, gs[1:])), bank


def _enable_dropout(blk, p, p_path):
    qb = blk.quad_block if hasattr(blk, "quad_block") else blk
    for m in qb.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = p
    qb.drop_path_rate = p_path
    return qb


CASES = [("qavitv2_c100", "blocks.0", False, 64, 64), ("hqavit_c100", "stage1_blocks.0", True, 64, 16),
         ("qavit_v1_c10", "blocks.1", False, 64, 64)]


@pytest.mark.parametrize("case,prefix,wrapped,ntok,nblk", CASES)
@pytest.mark.parametrize("p,p_path", [(0.1, 0.1), (0.25, 0.0), (0.0, 0.3)])
def test_block_dropout_fp32_vs_oracle_with_same_masks(case, prefix, wrapped, ntok, nblk, p, p_path):
    model, ocfg, sd, _ = build_model(case, precision="fp32")
    model.train()
    blk = model.get_submodule(prefix)
    _enable_dropout(blk, p, p_path)
    B = 3
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, ntok, 192, generator=g)
    wgt = torch.randn(B, ntok, 192, generator=g)
    masks = block_masks(SEED, 5, p, p_path, B, nblk, bf16=False)
    ref_out, ref_dx, ref_g, bank = _oracle_block(ocfg, sd, prefix, wrapped, x, wgt, masks)
    base_out = _oracle_block(ocfg, sd, prefix, wrapped, x, wgt, None)[0]
    assert rel_l2(ref_out, base_out) > 1e-2          # the masks do something
    st = _set_rng(SEED, 5)
    xl = x.cuda().requires_grad_(True)
    out = blk(xl)
    (out * wgt.cuda()).sum().backward()
    assert st.tolist() == [SEED, 6]                  # one snapshot per block call, advanced on the device
    assert rel_max(out, ref_out) < 1e-4
    assert rel_l2(xl.grad, ref_dx) < 1e-4
    named = dict(model.named_parameters())
    med = np.median([v.norm().item() for v in ref_g.values() if v is not None])
    for k, gr in ref_g.items():
        if gr is None:
            assert named[k].grad is None, k
            continue
        d = (named[k].grad.cpu() - gr).norm().item()
        tol = 3e-3 if gr.numel() <= 8 else 5e-4
        assert d <= tol * gr.norm().item() + 1e-5 * med, (k, d, gr.norm().item())
    bk, bv = bank.read()                             # the bank write sees the dropped branch output (H:465-468)
    assert rel_max(model.global_bank.global_k.data, bk) < 1e-5
    assert rel_max(model.global_bank.global_v.data, bv) < 1e-5


@pytest.mark.parametrize("case,prefix,wrapped,ntok,nblk", CASES[:2])
def test_block_dropout_bf16_vs_oracle_with_same_masks(case, prefix, wrapped, ntok, nblk):
    """bf16 run (tcgen05 GEMMs, mma.sync attention with the C-fragment mask layout) against the fp32 oracle fed the
    same masks: a wrong mask index shows up as O(p) = 1e-1 error, bf16 rounding as ~1e-2."""
    p, p_path = 0.1, 0.1
    model, ocfg, sd, _ = build_model(case, precision="bf16")
    model.train()
    blk = model.get_submodule(prefix)
    _enable_dropout(blk, p, p_path)
    B = 5
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, ntok, 192, generator=g)
    wgt = torch.randn(B, ntok, 192, generator=g)
    masks = block_masks(SEED, 0, p, p_path, B, nblk, bf16=True)
    ref_out, ref_dx, ref_g, bank = _oracle_block(ocfg, sd, prefix, wrapped, x, wgt, masks)
    nodrop_out, nodrop_dx, _, _ = _oracle_block(ocfg, sd, prefix, wrapped, x, wgt, None)
    _set_rng(SEED, 0)
    xl = x.cuda().requires_grad_(True)
    out = blk(xl)
    (out * wgt.cuda()).sum().backward()
    e_out, e_dx = rel_l2(out, ref_out), rel_l2(xl.grad, ref_dx)
    print(f"bf16 dropout block {case}: out {e_out:.3e} dx {e_dx:.3e}  (dropout effect itself: out {rel_l2(nodrop_out, ref_out):.3e} "
          f"dx {rel_l2(nodrop_dx, ref_dx):.3e})")
    assert e_out < 2e-2, e_out
    assert e_dx < 3e-2, e_dx
    named = dict(model.named_parameters())
    med = np.median([v.norm().item() for v in ref_g.values() if v is not None])
    zero_grads = ("token_upmix.upsample_attn.bias", "token_learner.attention.1.bias")
    worst = (0.0, "")
    for k, gr in ref_g.items():
        if gr is None or k.endswith(zero_grads):
            continue
        e = (named[k].grad.float().cpu() - gr).norm().item() / (gr.norm().item() + 5e-2 * med)
        worst = max(worst, (e, k))
    print("  worst parameter gradient", worst)
    assert worst[0] < 1.5e-1, worst


def test_dropout_fn_rate_replay_and_backward():
    import qavit_b200.functional as QF
    p = 0.1
    _set_rng(99, 0)
    x = torch.ones(1 << 22, device="cuda", requires_grad=True)
    y = QF.dropout(x, p, True)
    y.sum().backward()
    keep = 65536.0 / (65536.0 - 6554.0)
    zeros = (y == 0).float().mean().item()
    assert abs(zeros - 6554 / 65536) < 1e-3, zeros                     # 4 M draws: sigma = 1.5e-4
    assert torch.all((y == 0) | ((y - keep).abs() < 1e-6))
    assert torch.equal(x.grad, y.detach())                            # backward regenerates the same mask
    assert abs(y.mean().item() - 1.0) < 1e-3                          # unbiased
    y2 = QF.dropout(x.detach(), p, True)                              # offset advanced: a fresh mask
    assert (y2 != y).float().mean().item() > 0.1
    _set_rng(99, 0)
    assert torch.equal(QF.dropout(x.detach(), p, True), y.detach())   # same {seed, offset}: same mask
    assert QF.dropout(x, p, False) is x and QF.dropout(x, 0.0, True) is x


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_model_trains_with_reference_default_dropout(precision):
    """HQAViT with the reference's default dropout = drop_path = 0.1 (H:56-57) runs a train step with finite loss and
    gradients, draws a different mask on the next call, and ignores dropout in eval mode."""
    import qavit_b200 as Q
    torch.manual_seed(0)
    model = Q.HQAViT(Q.HQAViTConfig()).cuda().set_precision(precision)
    assert model.config.dropout == 0.1 and model.config.drop_path == 0.1
    x = torch.randn(8, 3, 32, 32, device="cuda")
    y = torch.randint(0, 100, (8,), device="cuda")
    model.train()
    _set_rng(7, 0)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
        l1 = model(x)
    loss = Q.cross_entropy(l1.float(), y, label_smoothing=0.1)
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
        l2 = model(x)
        assert rel_l2(l2.float(), l1.float()) > 1e-3                  # a different mask per call
        model.eval()
        e1, e2 = model(x), model(x)
        assert torch.equal(e1, e2)                                    # eval: no dropout


def test_hqavitv2_stem_droppath_matches_oracle():
    """DropPath of the HQAViTv2 stem's ConvNeXt blocks (HQAViTv2_CIFAR100.py:748, 787-799): per-image keep scales drawn in-kernel
    (site 0x5C00 + block, element = image) against the oracle fed the restated scales -- whole model, fp32 run, all gradients."""
    import qavit_b200 as Q
    from dropout_masks import Site
    model, ocfg, sd, _ = build_model("hqavitv2_c100", precision="fp32")
    model.train()
    rates = [0.0, 0.3, 0.0, 0.25, 0.1, 0.4, 0.2]
    for blk, r in zip(model.cnn_stem.blocks_in_order(), rates):
        blk.drop_path_rate = r
    B = 6
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, 32, 32, generator=g)
    y = torch.randint(0, 100, (B,), generator=g)
    seed, offset = 0x0BADC0FFEE123, 40
    object.__setattr__(model, "_stem_rng", torch.tensor([seed, offset], dtype=torch.int64, device="cuda"))
    logits = model(x.cuda())
    loss = Q.cross_entropy(logits, y.cuda(), label_smoothing=0.1)
    loss.backward()
    assert model._stem_rng.tolist()[1] > offset          # the state advanced on the device

    keeps = {name: Site(seed, offset, 0x5C00 + j, r).keep1(np.arange(B)) for j, (name, r) in enumerate(zip(O.V2_STEM_BLOCKS, rates)) if r > 0}
    assert any((k == 0).any() for k in keeps.values()), "pick a seed that drops something"

    def mask_fn(which, Bq, n):
        return keeps if which == "stem" else None
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd, ocfg, x, y, label_smoothing=0.1, mask_fn=mask_fn)
    assert rel_max(logits, ref_logits) < 1e-4
    assert abs(loss.item() - ref_loss.item()) < 2e-5
    named = dict(model.named_parameters())
    num = den = 0.0
    for k, gr in ref_grads.items():
        if gr is None:
            continue
        num += (named[k].grad.cpu() - gr).norm().item() ** 2
        den += gr.norm().item() ** 2
    assert (num / den) ** 0.5 < 1e-4
    for k in ("cnn_stem.stage3.1.gamma", "cnn_stem.stage4.0.pwconv2.weight", "cnn_stem.stem.0.weight", "cnn_stem.downsample2.0.weight"):
        assert rel_l2(named[k].grad, ref_grads[k]) < 5e-4, k


def test_splitfusion_dropout_matches_oracle():
    """SplitFusion's hard-coded Dropout(0.1) after cat_mlp's GELU (H:930): in-kernel masks (site 0x5F01, drop_rows ids on the
    [B N, d] matrix, one Philox offset per call) against the oracle fed the restated masks -- whole model, fp32 run, all gradients."""
    import qavit_b200 as Q
    from dropout_masks import Site
    model, ocfg, sd, _ = build_model("hqavit_c100", precision="fp32")
    model.train()
    ps = {"fuse2": 0.1, "fuse3": 0.3, "fuse4": 0.2}
    for n, p in ps.items():
        getattr(model, n).cat_mlp[3].p = p
    B, N, d = 5, 64, 192
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, 3, 32, 32, generator=g)
    y = torch.randint(0, 100, (B,), generator=g)
    seed, offset = 0x5EEDFACE1234, 100
    _set_rng(seed, offset)
    logits = model(x.cuda())
    loss = Q.cross_entropy(logits, y.cuda(), label_smoothing=0.1)
    loss.backward()
    # the three SplitFusion calls are the only rng users of this configuration: offsets offset, offset + 1, offset + 2
    keeps = {n: Site(seed, offset + i, 0x5F01, p).keep_rows(B * N, d).view(B, N, d) for i, (n, p) in enumerate(ps.items())}
    assert all(0.5 * p < (k == 0).float().mean().item() < 1.5 * p for (n, p), k in zip(ps.items(), keeps.values()))

    def mask_fn(which, Bq, n):
        return keeps.get(which) if isinstance(which, str) else None
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd, ocfg, x, y, label_smoothing=0.1, mask_fn=mask_fn)
    assert rel_max(logits, ref_logits) < 1e-4
    assert abs(loss.item() - ref_loss.item()) < 2e-5
    named = dict(model.named_parameters())
    num = den = 0.0
    for k, gr in ref_grads.items():
        if gr is None:
            continue
        num += (named[k].grad.cpu() - gr).norm().item() ** 2
        den += gr.norm().item() ** 2
    assert (num / den) ** 0.5 < 1e-4
