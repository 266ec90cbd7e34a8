"""Pins the CPU oracle (oracle/qavit_oracle.py) against the golden vectors produced by the
live reference (tests/golden/make_golden.py) and, when /root/reference is mounted, against
the reference modules directly.  CPU only."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import qavit_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)

CASES = list(MG.CASES)
FAST = ["hqavit_c100", "qavit_v1_c10"]


def golden(case):
    return np.load(os.path.join(HERE, "golden", case + ".npz"))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)


def ocfg_of(case):
    return O.OracleConfig(**MG.CASES[case][4]), MG.CASES[case][5]


@pytest.mark.parametrize("case", CASES)
def test_eval_logits_match_reference_golden(case):
    cfg, B = ocfg_of(case)
    g = golden(case)
    sd = O.synthetic_state(cfg)
    x, _ = MG.inputs(cfg, B)
    with torch.no_grad():
        lo = O.forward(sd, cfg, x, train=False).numpy()
    assert rel(lo, g["eval_logits"]) < 2e-5
    assert (lo.argmax(-1) == g["eval_logits"].argmax(-1)).all()


@pytest.mark.parametrize("case", CASES)
def test_train_step_matches_reference_golden(case):
    cfg, B = ocfg_of(case)
    g = golden(case)
    sd = O.synthetic_state(cfg)
    x, y = MG.inputs(cfg, B)
    logits, loss, grads, new_state = O.loss_and_grads(sd, cfg, x, y, label_smoothing=0.1)
    assert rel(logits.numpy(), g["train_logits"]) < 2e-5
    assert abs(float(loss) - float(g["train_loss"])) < 1e-5
    assert rel(new_state["global_bank.global_k"].numpy(), g["bank_k_after_fwd"]) < 1e-6
    assert rel(new_state["global_bank.global_v"].numpy(), g["bank_v_after_fwd"]) < 1e-6
    if "update_count_after_fwd" in g:
        assert int(new_state["global_bank.update_count"]) == int(g["update_count_after_fwd"])
    keys = list(g["keys"])
    assert keys == O.trainable_keys(cfg)
    fp = np.stack([MG.fingerprint(grads[k]) for k in keys])
    ref = g["grad_fp"]
    none_ref = np.isnan(ref[:, 0])
    assert (np.isnan(fp[:, 0]) == none_ref).all(), "set of grad-is-None params differs"
    # Per-parameter gradient L2 norm and leading values: |delta| <= tol * that parameter's norm + 1e-6 * the
    # median norm, tol = 5e-4 (3e-3 for parameters of <= 8 elements).  Measured here against an fp64 run of the
    # reference: the reference's OWN fp32 gradients are off by up to 7.6e-4 on scalar parameters (gamma,
    # fusion_weights, beta: cancellation-heavy sums) and ~1e-4 elsewhere, and its mathematically-zero gradients
    # (biases feeding a softmax over the same axis or a LayerNorm/BatchNorm) are pure noise ~1e-8 -- hence the
    # absolute term.  In fp64 the oracle equals the reference to 1e-12 (test below).  The aggregate over all
    # parameters is held to 1e-4.
    med = np.nanmedian(ref[:, 0])
    r0, f0 = ref[~none_ref], fp[~none_ref]
    tol = np.where(r0[:, 2] <= 8, 3e-3, 5e-4)
    bad = np.abs(f0[:, 0] - r0[:, 0]) > tol * r0[:, 0] + 1e-6 * med
    assert not bad.any(), [keys[i] for i in np.flatnonzero(~none_ref)[bad]]
    assert (np.abs(f0[:, 3:] - r0[:, 3:]).max(1) <= tol * r0[:, 0] + 1e-6 * med).all()
    tot_ref = np.sqrt((r0[:, 0] ** 2).sum())
    assert abs(np.sqrt((f0[:, 0] ** 2).sum()) - tot_ref) < 1e-4 * tot_ref


@pytest.mark.parametrize("case", FAST)
def test_three_optimizer_steps_match_reference_golden(case):
    cfg, B = ocfg_of(case)
    g = golden(case)
    sd = O.synthetic_state(cfg)
    opt_state = {}
    keys = O.trainable_keys(cfg)
    for s in range(3):
        x, y = MG.inputs(cfg, B, seed=1234 + s)
        _, loss, grads, new_state = O.loss_and_grads(sd, cfg, x, y, label_smoothing=0.1)
        assert abs(float(loss) - g["step_losses"][s]) < 2e-5
        sd.update(new_state)
        O.clip_grads_(grads)
        O.adamw_step_({k: sd[k] for k in keys}, grads, opt_state, lr=6e-4, wd=0.06)
    fp = np.stack([MG.fingerprint(sd[k]) for k in keys])
    # Adam normalises the update, so a parameter whose true gradient is zero (fp32 noise ~1e-8, see above)
    # still moves by ~lr per step in a noise-determined direction: those get 3*lr, the rest 2e-5 absolute.
    gn = g["grad_fp"][:, 0]
    noise = np.isnan(gn) | (gn < 1e-5 * np.nanmedian(gn))
    d = np.abs(fp[:, 3:] - g["param_fp_after_steps"][:, 3:]).max(1)
    assert d[~noise].max() < 2e-5
    assert d[noise].max() < 3 * 6e-4 + 2e-5
    x, _ = MG.inputs(cfg, B)
    with torch.no_grad():
        lo = O.forward(sd, cfg, x, train=False).numpy()
    # 2e-3, not 1e-4: the conv biases in front of BatchNorm are among those noise-driven parameters; their
    # drift cancels in train mode (batch statistics) but not in eval mode (running statistics).
    assert rel(lo, g["eval_logits_after_steps"]) < 2e-3


@pytest.mark.skipif(not os.path.isdir(MG.REF), reason="live reference not mounted (GPU box)")
@pytest.mark.parametrize("case", ["hqavit_c100", "qavitv2_c100", "qavit_v1_224", "hqavitv2_c100"])
def test_oracle_equals_live_reference_in_fp64(case):
    """In double precision the restatement and the reference agree to rounding on every gradient element."""
    model, cfg, B = MG.build_reference(case)
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in O.synthetic_state(cfg).items()}
    model.double()
    model.load_state_dict(O.with_bank_aliases(sd, cfg), strict=True)
    x, y = MG.inputs(cfg, B)
    x = x.double()
    model.train()
    loss = torch.nn.functional.cross_entropy(model(x), y, label_smoothing=0.1)
    loss.backward()
    _, oloss, grads, new_state = O.loss_and_grads(sd, cfg, x, y, label_smoothing=0.1)
    assert abs(float(loss) - float(oloss)) < 1e-12
    assert (new_state["global_bank.global_k"] - model.global_bank.global_k).abs().max() < 1e-14
    gmed = np.median([p.grad.norm().item() for p in model.parameters() if p.grad is not None])
    for n, p in model.named_parameters():
        if p.grad is None:
            assert grads[n] is None, n
        else:
            assert (grads[n] - p.grad).norm().item() <= 1e-10 * p.grad.norm().item() + 1e-12 * gmed, n


# ----------------------------------------------------------------------------------------------- dropout placement
@pytest.mark.skipif(not os.path.isdir(MG.REF), reason="live reference not mounted (GPU box)")
@pytest.mark.parametrize("case,prefix,wrapped,ntok,nblk", [("qavitv2_c100", "blocks.3", False, 64, 64),
                                                          ("hqavit_c100", "stage2_blocks.1", True, 64, 16)])
def test_oracle_dropout_sites_match_live_reference(case, prefix, wrapped, ntok, nblk, monkeypatch):
    """Pins WHERE the oracle applies dropout / DropPath (quad_block's `masks`) to the live reference block in train mode:
    the reference's three random sources -- SDPA dropout_p, nn.Dropout, drop_path() -- are replaced by explicit masks
    handed out in call order, the oracle gets the same masks by name, and in fp64 the two agree to rounding on the
    output, dx, every parameter gradient and the bank state the (dropped) branch outputs are written into."""
    mod_name, cls, cfg_cls, over, okw, _ = MG.CASES[case]
    mod = MG.import_reference(mod_name)
    mod.HAS_FLASH_ATTN = False
    p, p_path = 0.1, 0.2
    torch.manual_seed(3)
    model = getattr(mod, cls)(getattr(mod, cfg_cls)(dropout=p, drop_path=p_path, **over))
    cfg = O.OracleConfig(**okw)
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in O.synthetic_state(cfg).items()}
    model.double()
    model.load_state_dict(O.with_bank_aliases(sd, cfg), strict=True)
    model.train()
    blk = model.get_submodule(prefix)
    B = 3
    g = torch.Generator().manual_seed(5)
    masks = {k: v.double() for k, v in O.random_masks(cfg, B, nblk, p, p_path, g).items()}
    x = torch.randn(B, ntok, cfg.embed_dim, generator=g, dtype=torch.float64)
    wgt = torch.randn(B, ntok, cfg.embed_dim, generator=g, dtype=torch.float64)

    # proj_swa is applied by the reference BEFORE the window reverse (H:465-466): hand it over in window layout
    s, w = int(nblk ** 0.5), cfg.window_size
    pw = masks["proj_swa"].view(B, s // w, w, s // w, w, -1).permute(0, 1, 3, 2, 4, 5).reshape(-1, w * w, cfg.embed_dim)
    q_att = [masks[k] for k in ("att_swa", "att_msda", "att_cga", "att_cross")]
    q_drop = [pw, masks["proj_msda"], masks["proj_cga"], masks["proj_cross"], masks["b1"], masks["b2"], masks["ffn"]]
    q_path = [masks["path1"], masks["path2"]]

    def sdpa(q, k, v, dropout_p=0.0, **kw):
        assert dropout_p == p
        P = torch.softmax((q @ k.transpose(-1, -2)) / q.shape[-1] ** 0.5, -1)
        return (P * q_att.pop(0)) @ v

    def dropout_forward(self, t):
        if not self.training or self.p == 0:
            return t
        assert self.p == p
        return t * q_drop.pop(0)

    def drop_path(t, drop_prob=0.0, training=False):
        if drop_prob == 0.0 or not training:
            return t
        return t * q_path.pop(0)

    monkeypatch.setattr(torch.nn.functional, "scaled_dot_product_attention", sdpa)
    monkeypatch.setattr(torch.nn.Dropout, "forward", dropout_forward)
    monkeypatch.setattr(mod, "drop_path", drop_path)
    qb = blk.quad_block if wrapped else blk
    assert 0 < qb.drop_path1.drop_prob < p_path + 1e-9           # H:1187: linspace(0, drop_path, depth)
    xl = x.clone().requires_grad_(True)
    out = blk(xl)
    (out * wgt).sum().backward()
    assert not q_att and not q_drop and not q_path               # every site was visited exactly once

    keys = [k for k in O.trainable_keys(cfg) if k.startswith(prefix + ".") or k.startswith("global_bank.")]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    full = dict(sd)
    full.update(leaves)
    xo = x.clone().requires_grad_(True)
    bank = O.Bank(full, cfg)
    oo = (O.wrapped_block if wrapped else O.quad_block)(xo, full, prefix, cfg, bank, True, masks)
    gs = torch.autograd.grad((oo * wgt).sum(), [xo] + list(leaves.values()), allow_unused=True)
    assert (oo - out).abs().max() < 1e-11
    assert (gs[0] - xl.grad).abs().max() < 1e-10 * xl.grad.abs().max()
    named = dict(model.named_parameters())
    for k, gr in zip(keys, gs[1:]):
        if gr is None:
            assert named[k].grad is None, k
        else:
            assert (gr - named[k].grad).norm() <= 1e-9 * named[k].grad.norm() + 1e-12, k
    bk, bv = bank.read()
    assert (bk - model.global_bank.global_k).abs().max() < 1e-13
    assert (bv - model.global_bank.global_v).abs().max() < 1e-13
