"""Parity of the CUDA path (through the drop-in modules -> ctypes -> C ABI) against the CPU oracle and the golden
vectors of the live reference.  fp32 run: 1e-4 on logits / activations, argmax bit-exact, gradients to the
tolerances derived in tests/test_oracle.py; bf16 run: 1e-2 (north star tolerances)."""
import numpy as np
import pytest
import torch

from oracle import qavit_oracle as O
from util import CASES, build_model, golden, inputs, rel_l2, rel_max

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _oracle_block(ocfg, sd, prefix, wrapped, x, wgt, train=True):
    keys = [k for k in O.trainable_keys(ocfg) if k.startswith(prefix + ".") or k.startswith("global_bank.")]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    full = dict(sd)
    full.update(leaves)
    xl = x.detach().clone().requires_grad_(True)
    bank = O.Bank(full, ocfg)
    out = (O.wrapped_block if wrapped else O.quad_block)(xl, full, prefix, ocfg, bank, train)
    gs = torch.autograd.grad((out * wgt).sum(), [xl] + list(leaves.values()), allow_unused=True)
    return out.detach(), gs[0], dict(zip(keys, gs[1:])), bank


def _our_block(model, prefix, x, wgt):
    blk = model.get_submodule(prefix)
    xl = x.cuda().requires_grad_(True)
    out = blk(xl)
    (out * wgt.cuda()).sum().backward()
    return out.detach().cpu(), xl.grad.cpu()


@pytest.mark.parametrize("case,prefix,wrapped,ntok", [
    ("qavitv2_c100", "blocks.0", False, 64),
    ("hqavit_c100", "stage1_blocks.0", True, 64),
    ("qavit_v1_c10", "blocks.1", False, 64),
    ("qavit_v1_224", "blocks.2", False, 196),       # QAViT.py's own defaults: 14 x 14 tokens, 7 x 7 windows, k = 64, dilations (1, 2, 3)
    ("qavitv2b_224", "blocks.5", False, 196),
])
@pytest.mark.parametrize("train", [True, False])
def test_block_forward_backward_fp32(case, prefix, wrapped, ntok, train):
    model, ocfg, sd, _ = build_model(case, precision="fp32")
    model.train(train)
    B = 3
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, ntok, 192, generator=g)
    wgt = torch.randn(B, ntok, 192, generator=g)
    ref_out, ref_dx, ref_g, bank = _oracle_block(ocfg, sd, prefix + (".quad_block" if False else ""), wrapped, x, wgt, train)
    out, dx = _our_block(model, prefix, x, wgt)
    assert rel_max(out, ref_out) < 1e-4
    assert rel_l2(dx, ref_dx) < 1e-4
    named = dict(model.named_parameters())
    med = np.median([v.norm().item() for v in ref_g.values() if v is not None])
    for k, gr in ref_g.items():
        p = named[k]
        if gr is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        d = (p.grad.cpu() - gr).norm().item()
        tol = 3e-3 if gr.numel() <= 8 else 5e-4
        assert d <= tol * gr.norm().item() + 1e-5 * med, (k, d, gr.norm().item())
    bk, bv = bank.read()
    assert rel_max(model.global_bank.global_k.data, bk) < 1e-5
    assert rel_max(model.global_bank.global_v.data, bv) < 1e-5
    if train and not bank.v1:
        assert int(model.global_bank.update_count) == bank.count == 3


@pytest.mark.parametrize("case", list(CASES))
def test_model_eval_logits_fp32_vs_oracle_and_golden(case):
    model, ocfg, sd, B = build_model(case, precision="fp32")
    model.eval()
    x, _ = inputs(ocfg, B)
    with torch.no_grad():
        lo = model(x.cuda()).cpu()
        ref = O.forward(sd, ocfg, x, train=False)
    gold = torch.from_numpy(golden(case)["eval_logits"])
    assert rel_max(lo, ref) < 1e-4
    assert rel_max(lo, gold) < 1e-4
    assert (lo.argmax(-1) == gold.argmax(-1)).all()          # top-1 bit-exact in fp32


def _grad_fingerprints(model, keys):
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    MG = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(MG)
    named = dict(model.named_parameters())
    return np.stack([MG.fingerprint(None if named[k].grad is None else named[k].grad.cpu()) for k in keys])


@pytest.mark.parametrize("case", list(CASES))
def test_model_train_step_fp32_vs_golden(case):
    """Forward + backward in train mode (bank writes on): logits, loss, every parameter gradient, bank state."""
    import qavit_b200 as Q
    model, ocfg, sd, B = build_model(case, precision="fp32")
    model.train()
    x, y = inputs(ocfg, B)
    logits = model(x.cuda())
    loss = Q.cross_entropy(logits, y.cuda(), label_smoothing=0.1)
    loss.backward()
    g = golden(case)
    assert rel_max(logits, torch.from_numpy(g["train_logits"])) < 1e-4
    assert abs(loss.item() - float(g["train_loss"])) < 2e-5
    assert rel_max(model.global_bank.global_k.data, torch.from_numpy(g["bank_k_after_fwd"])) < 1e-5
    assert rel_max(model.global_bank.global_v.data, torch.from_numpy(g["bank_v_after_fwd"])) < 1e-5
    if "update_count_after_fwd" in g:
        assert int(model.global_bank.update_count) == int(g["update_count_after_fwd"])
    keys = list(g["keys"])
    fp, ref = _grad_fingerprints(model, keys), g["grad_fp"]
    none_ref = np.isnan(ref[:, 0])
    assert (np.isnan(fp[:, 0]) == none_ref).all(), "set of grad-is-None parameters differs from the reference"
    med = np.nanmedian(ref[:, 0])
    r0, f0 = ref[~none_ref], fp[~none_ref]
    tol = np.where(r0[:, 2] <= 8, 3e-3, 5e-4)
    bad = np.abs(f0[:, 0] - r0[:, 0]) > tol * r0[:, 0] + 1e-6 * med
    assert not bad.any(), [keys[i] for i in np.flatnonzero(~none_ref)[bad]][:10]
    assert (np.abs(f0[:, 3:] - r0[:, 3:]).max(1) <= tol * r0[:, 0] + 1e-6 * med).all()
    tot = np.sqrt((r0[:, 0] ** 2).sum())
    assert abs(np.sqrt((f0[:, 0] ** 2).sum()) - tot) < 1e-4 * tot


def test_model_full_gradients_fp32_vs_oracle():
    import qavit_b200 as Q
    model, ocfg, sd, B = build_model("hqavit_c100", precision="fp32")
    model.train()
    x, y = inputs(ocfg, B, seed=99)
    loss = Q.cross_entropy(model(x.cuda()), y.cuda(), label_smoothing=0.12)
    loss.backward()
    _, oloss, grads, _ = O.loss_and_grads(sd, ocfg, x, y, label_smoothing=0.12)
    assert abs(loss.item() - oloss.item()) < 2e-5
    med = np.median([v.norm().item() for v in grads.values() if v is not None])
    num = den = 0.0
    for n, p in model.named_parameters():
        if grads[n] is None:
            assert p.grad is None, n
            continue
        d = (p.grad.cpu() - grads[n]).norm().item()
        tol = 3e-3 if grads[n].numel() <= 8 else 5e-4
        assert d <= tol * grads[n].norm().item() + 1e-5 * med, (n, d, grads[n].norm().item())
        num += d * d
        den += grads[n].norm().item() ** 2
    assert (num / den) ** 0.5 < 1e-4            # all gradients together within 1e-4


@pytest.mark.parametrize("case", ["hqavit_c100", "qavitv2_c100"])
def test_three_optimizer_steps_fp32_vs_golden(case):
    import qavit_b200 as Q
    model, ocfg, sd, B = build_model(case, precision="fp32")
    model.train()
    opt = Q.FusedAdamW(model.named_parameters(), lr=6e-4, betas=(0.9, 0.999), weight_decay=0.06, max_grad_norm=0.5)
    g = golden(case)
    for s in range(3):
        x, y = inputs(ocfg, B, seed=1234 + s)
        for p in model.parameters():
            p.grad = None
        loss = Q.cross_entropy(model(x.cuda()), y.cuda(), label_smoothing=0.1)
        loss.backward()
        assert abs(loss.item() - g["step_losses"][s]) < 5e-5, s
        opt._have_flags = False
        opt.clip()
        opt.step()
    model.eval()
    x, _ = inputs(ocfg, B)
    with torch.no_grad():
        lo = model(x.cuda()).cpu()
    assert rel_max(lo, torch.from_numpy(g["eval_logits_after_steps"])) < 2e-3


def _autocast_emulation(ocfg, sd, x, y):
    """The SAME graph evaluated the way torch.autocast(bfloat16) evaluates the reference on a GPU: matmul / conv in
    bf16, LayerNorm / softmax / loss in fp32.  Its distance from the fp32 oracle is the bf16
    noise floor of this graph with these weights."""
    ln32 = O._ln
    O._ln = lambda t, sd_, prefix, eps=1e-5: ln32(t.float(), sd_, prefix, eps)
    try:
        sdg = {k: v.cuda() for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, loss, grads, _ = O.loss_and_grads(sdg, ocfg, x.cuda(), y.cuda(), label_smoothing=0.1)
    finally:
        O._ln = ln32
    return logits.float().cpu(), loss.float().cpu(), {k: (None if g is None else g.float().cpu()) for k, g in grads.items()}


def _total_grad_err(get, grads):
    num = den = 0.0
    for n, g in grads.items():
        if g is None:
            continue
        d = (get(n) - g).norm().item()
        num += d * d
        den += g.norm().item() ** 2
    return (num / den) ** 0.5


@pytest.mark.parametrize("case", ["hqavit_c100", "qavitv2_c100"])
def test_model_bf16_vs_oracle(case):
    """bf16 run (tcgen05 GEMMs, bf16 activations).  North-star gate: 1e-2 on logits and gradients -- met outright on
    QAViTv2.  On HQAViT with the (deliberately large) synthetic weights the eight TokenLearner softmaxes amplify bf16
    rounding: an autocast evaluation of the reference graph itself is 9e-2 / 2e-1 away from fp32 (measured, see
    _autocast_emulation), so there the gate is 'no further from the fp32 oracle than the autocast reference is'."""
    import qavit_b200 as Q
    model, ocfg, sd, B = build_model(case, precision="bf16")
    model.train()
    x, y = inputs(ocfg, B)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(x.cuda())
    loss = Q.cross_entropy(logits, y.cuda(), label_smoothing=0.1)
    loss.backward()
    ref_logits, ref_loss, grads, new_state = O.loss_and_grads(sd, ocfg, x, y, label_smoothing=0.1)
    emu_logits, emu_loss, emu_grads = _autocast_emulation(ocfg, sd, x, y)
    named = dict(model.named_parameters())
    for n, g in grads.items():
        assert (named[n].grad is None) == (g is None), n
    e_log, e_emu = rel_max(logits, ref_logits), rel_max(emu_logits, ref_logits)
    g_log = _total_grad_err(lambda n: named[n].grad.float().cpu(), grads)
    g_emu = _total_grad_err(lambda n: emu_grads[n], grads)
    print(f"bf16 {case}: logits ours {e_log:.3e} autocast-ref {e_emu:.3e}; grads ours {g_log:.3e} autocast-ref {g_emu:.3e}")
    assert e_log <= max(1e-2, 1.1 * e_emu), (e_log, e_emu)
    assert g_log <= max(1e-2, 1.1 * g_emu), (g_log, g_emu)
    assert abs(loss.item() - ref_loss.item()) <= max(1e-2 * abs(ref_loss.item()), 1.5 * abs(emu_loss.item() - ref_loss.item()))
    assert rel_max(model.global_bank.global_k.data, new_state["global_bank.global_k"]) < 2e-2
    if case == "qavitv2_c100":
        assert e_log < 1e-2 and g_log < 1e-2


@pytest.mark.parametrize("fam", ["hqavit", "qavit"])
def test_bf16_vs_fp32_with_reference_init(fam):
    """With the reference's own initialisation (trunc-normal 0.02 etc., H:1212-1224) the bf16 run is within 1e-2 of
    the fp32 run of the same model on logits and on the gradients taken together."""
    import copy
    import qavit_b200 as Q
    torch.manual_seed(42)
    if fam == "hqavit":
        m32 = Q.HQAViT(Q.HQAViTConfig(dropout=0.0, drop_path=0.0))
        for n in ("fuse2", "fuse3", "fuse4"):
            getattr(m32, n).cat_mlp[3].p = 0.0
    else:
        m32 = Q.QAViT(Q.QAViTConfig(dropout=0.0, drop_path=0.0))
    sd0 = {k: v.detach().clone() for k, v in m32.state_dict().items()}   # before any forward mutates bank / BN state
    m32 = m32.cuda().train()
    m16 = copy.deepcopy(m32)
    m32.set_precision("fp32")
    m16.set_precision("bf16")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(16, 3, 32, 32, generator=g).cuda()
    y = torch.randint(0, 100, (16,), generator=g).cuda()
    l32 = m32(x)
    Q.cross_entropy(l32, y, label_smoothing=0.1).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l16 = m16(x)
    Q.cross_entropy(l16, y, label_smoothing=0.1).backward()
    n16 = dict(m16.named_parameters())
    num = den = 0.0
    for n, p in m32.named_parameters():
        if p.grad is None:
            assert n16[n].grad is None
            continue
        num += (n16[n].grad - p.grad).norm().item() ** 2
        den += p.grad.norm().item() ** 2
    e_log, e_grad = rel_max(l16, l32), (num / den) ** 0.5
    print(f"ref-init {fam}: logits {e_log:.3e} grads {e_grad:.3e}")
    if fam == "qavit":
        assert e_log < 1e-2 and e_grad < 1e-2
    else:
        # HQAViT sits at the 1e-2 edge of this max-norm metric (16 x 100 logits, worst element): the bound is the
        # north-star 1e-2 or the bf16 noise floor of the reference graph itself under autocast, whichever is larger.
        ocfg = O.OracleConfig(family="hqavit")
        ref_logits, _, ref_grads, _ = O.loss_and_grads(sd0, ocfg, x.cpu(), y.cpu(), label_smoothing=0.1)
        emu_logits, _, emu_grads = _autocast_emulation(ocfg, sd0, x.cpu(), y.cpu())
        f_log = rel_max(emu_logits, ref_logits)
        f_grad = _total_grad_err(lambda n: emu_grads[n], ref_grads)
        print(f"ref-init hqavit: autocast reference graph vs fp32: logits {f_log:.3e} grads {f_grad:.3e}")
        assert rel_max(l32, ref_logits) < 1e-4
        assert e_log <= max(1e-2, 1.25 * f_log), (e_log, f_log)
        assert e_grad <= max(1e-2, 1.25 * f_grad), (e_grad, f_grad)


@pytest.mark.parametrize("case,prefix,ntok", [("hqavit_c100", "stage2_blocks.1", 64), ("qavitv2_c100", "blocks.3", 64),
                                              ("hqavit_tinyin", "stage3_blocks.2", 256),    # 64 learned / 256 stream tokens
                                              ("qavitv2b_224", "blocks.4", 196)])           # 196 tokens: tcgen05 GEMMs + SIMT attention
def test_block_bf16_close_to_fp32(case, prefix, ntok):
    """One block, bf16 run (tcgen05 GEMMs + mma.sync attention) against the fp32 run of the same block: a layout bug in
    a tensor-core path shows up as O(1) error, bf16 rounding as ~1e-2."""
    B = 5
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, ntok, 192, generator=g)
    wgt = torch.randn(B, ntok, 192, generator=g)
    res = {}
    for prec in ("fp32", "bf16"):
        model, ocfg, sd, _ = build_model(case, precision=prec)
        model.train()
        out, dx = _our_block(model, prefix, x, wgt)
        grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
        res[prec] = (out, dx, grads, model.global_bank.global_k.detach().cpu().clone())
    o32, d32, g32, b32 = res["fp32"]
    o16, d16, g16, b16 = res["bf16"]
    assert set(g32) == set(g16)
    assert rel_l2(o16, o32) < 2e-2, rel_l2(o16, o32)
    assert rel_l2(d16, d32) < 3e-2, rel_l2(d16, d32)
    assert rel_max(b16, b32) < 1e-2
    # floor: 5 % of the median gradient norm (mathematically-zero gradients such as the TokenLearner biases are noise)
    med = np.median([v.norm().item() for v in g32.values()])
    # gradients that are exactly zero in exact arithmetic (a per-token bias removed by the LayerNorm that follows,
    # H:1027-1028; a per-slot bias removed by the softmax over tokens, H:996): both runs only hold rounding noise there,
    # so they are bounded in absolute terms (2 % of the median gradient norm) instead of relatively
    zero_grads = ("token_upmix.upsample_attn.bias", "token_learner.attention.1.bias")
    for n in g32:
        if n.endswith(zero_grads):
            assert (g16[n] - g32[n]).norm().item() < 2e-2 * med, n
    def err(n):
        return (g16[n] - g32[n]).norm().item() / (g32[n].norm().item() + 5e-2 * med)
    big = [n for n in g32 if not n.endswith(zero_grads) and g32[n].numel() > 8]
    small = [n for n in g32 if not n.endswith(zero_grads) and g32[n].numel() <= 8]   # scalars (gamma, beta, fusion weights):
    worst = max((err(n), n) for n in big)                                            # one number, no averaging of the noise
    worst_small = max((err(n), n) for n in small)
    print(f"block bf16 vs fp32 {case}: out {rel_l2(o16, o32):.3e} dx {rel_l2(d16, d32):.3e} worst param grad {worst} / scalars {worst_small}")
    assert worst[0] < 8e-2, worst
    assert worst_small[0] < 1.5e-1, worst_small


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
