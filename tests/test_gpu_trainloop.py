"""The pieces of the reference's train loop that sit around the model step: ModelEMA folded into the
AdamW kernel, GradientMonitor's per-tensor norms, on-device CutMix / MixUp -- each against the reference's torch expression."""
import types

import numpy as np
import pytest
import torch

import qavit_b200 as Q
from util import build_model

pytestmark = pytest.mark.gpu


def _toy_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(192, 192), (192,), (1,), (16, 48), (3, 5, 7), (100, 192)]
    return [(f"p{i}.{'dwconv.w' if i == 3 else 'w'}", torch.nn.Parameter(torch.randn(s, generator=g).cuda())) for i, s in enumerate(shapes)]


def test_adamw_with_fused_ema_matches_torch_adamw_plus_model_ema_update():
    named = _toy_params()
    ref = [p.detach().clone().requires_grad_(True) for _, p in named]
    opt = Q.FusedAdamW(named, lr=3e-3, betas=(0.9, 0.999), weight_decay=0.05, max_grad_norm=None)
    ema = opt.enable_ema(0.99)
    topt = torch.optim.AdamW(ref, lr=3e-3, betas=(0.9, 0.999), weight_decay=0.05)
    tema = [p.detach().clone() for p in ref]
    has_grad = [True, True, False, True, True, True]        # p2 never receives a gradient (like the bank's write_* tensors)
    opt.set_grad_mask(has_grad)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        if step == 2:
            opt.ema_decay = 0.9                               # ModelEMA.set_decay during warm-up (H:1633-1638)
        opt.zero_grad()
        for (n, p), r, h in zip(named, ref, has_grad):
            gr = torch.randn(p.shape, generator=g).cuda()
            if h:
                p.grad.copy_(gr)
                r.grad = gr.clone()
            else:
                r.grad = None
        opt.step()
        topt.step()
        d = opt.ema_decay
        with torch.no_grad():
            for e, r in zip(tema, ref):                       # ModelEMA.update, H:146-149
                e.mul_(d).add_(r.detach(), alpha=1.0 - d)
    views = opt.ema_views()
    for (n, p), r, e in zip(named, ref, tema):
        assert torch.allclose(p.detach(), r.detach(), rtol=2e-6, atol=1e-7), n
        assert torch.allclose(views[n], e, rtol=2e-6, atol=1e-7), n
    assert ema.data_ptr() == opt.flat_ema.data_ptr()


def test_monitor_norms_match_torch():
    named = _toy_params(3)
    opt = Q.FusedAdamW(named, lr=1e-3)
    opt.zero_grad()
    for _, p in named:
        p.grad.copy_(torch.randn_like(p))
    gn, pn = opt.monitor_norms()
    for i, (n, p) in enumerate(named):
        assert abs(gn[i].item() - p.grad.norm().item()) <= 1e-5 * p.grad.norm().item() + 1e-7, n
        assert abs(pn[i].item() - p.norm().item()) <= 1e-5 * p.norm().item() + 1e-7, n


def test_model_ema_drop_in_tracks_parameters_and_buffers():
    model, ocfg, sd, B = build_model("qavitv2_c100", precision="fp32")
    model.train()
    opt = Q.FusedAdamW(model.named_parameters(), lr=1e-2, weight_decay=0.0, max_grad_norm=0.5)
    ema = Q.ModelEMA(model, decay=0.9, optimizer=opt)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    x = torch.randn(2, 3, 32, 32, device="cuda")
    y = torch.randint(0, 100, (2,), device="cuda")
    opt.zero_grad()
    Q.cross_entropy(model(x), y).backward()
    opt.clip()
    opt.step()
    ema.update(model)
    ema_p = dict(ema.ema.named_parameters())
    moved = 0
    for n, p in model.named_parameters():
        want = 0.9 * before[n] + 0.1 * p.detach()
        assert torch.allclose(ema_p[n], want, rtol=1e-5, atol=1e-7), n
        moved += int(not torch.equal(p.detach(), before[n]))
    assert moved > 500
    assert int(ema.ema.global_bank.update_count) == int(model.global_bank.update_count) > 0   # buffers mirrored (H:151-156)
    assert not ema.ema.training
    pd, bd = ema.compute_distance(model)
    assert pd > 0 and bd == 0.0
    with torch.no_grad():
        assert ema.ema(x).shape == (2, 100)                  # the averaged copy is a working model


@pytest.mark.parametrize("shape", [(7, 3, 32, 32), (4, 3, 64, 64)])
def test_batch_mix_matches_reference_expressions(shape):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g).cuda()
    perm = torch.randperm(shape[0], generator=g).cuda()
    x1, y1, x2, y2 = 3, 0, 20, 17
    want = x.clone()
    want[:, :, y1:y2, x1:x2] = x[perm, :, y1:y2, x1:x2]      # H:1386
    assert torch.equal(Q.batch_mix(x, perm, "cutmix", box=(x1, y1, x2, y2)), want)
    lam = 0.37219
    assert torch.equal(Q.batch_mix(x, perm, "mixup", lam=lam), lam * x + (1 - lam) * x[perm])   # H:1396, bit-exact


def test_mix_batch_draw_order_and_two_target_loss():
    cfg = types.SimpleNamespace(use_cutmix=True, cutmix_alpha=1.0, use_mixup=True, mixup_alpha=0.2, mix_prob=0.6)
    x = torch.randn(16, 3, 32, 32, device="cuda")
    y = torch.randint(0, 100, (16,), device="cuda")
    seen = set()
    for seed in range(12):
        rng = np.random.RandomState(seed)
        probe = np.random.RandomState(seed)
        first = probe.rand()
        out, ta, tb, lam, kind = Q.mix_batch(x, y, cfg, rng)
        seen.add(kind)
        if first < cfg.mix_prob:                              # the first draw decides CutMix (H:1382)
            assert kind == "cutmix" and 0.0 <= lam <= 1.0
            assert torch.equal(ta, y) and out.shape == x.shape
        elif kind is None:
            assert out is x and lam == 1.0
        else:
            assert kind == "mixup"
    assert "cutmix" in seen and (None in seen or "mixup" in seen)
    logits = torch.randn(16, 100, device="cuda", requires_grad=True)
    out, ta, tb, lam, kind = Q.mix_batch(x, y, cfg, np.random.RandomState(0))
    loss = Q.cross_entropy(logits, ta, label_smoothing=0.1, target_b=tb, lam=lam)
    ce = torch.nn.CrossEntropyLoss(label_smoothing=0.1)
    want = lam * ce(logits, ta) + (1.0 - lam) * ce(logits, tb)     # H:1406
    assert abs(loss.item() - want.item()) < 1e-5


def test_gradient_monitor_drop_in_matches_reference_algorithm():
    """GradientMonitor.log_gradients (H:197-242) restated with per-tensor torch norms vs the two-launch version."""
    named = _toy_params(5)
    opt = Q.FusedAdamW(named, lr=1e-3)
    has_grad = [True, True, False, True, True, True]
    opt.set_grad_mask(has_grad)
    opt.zero_grad()
    g = torch.Generator().manual_seed(9)
    for i, (_, p) in enumerate(named):
        p.grad.copy_(torch.randn(p.shape, generator=g).cuda() * (40.0 if i == 3 else 1.0))   # p3: norm > 10 -> detailed stats
    mon = Q.GradientMonitor(optimizer=opt)
    total, pnorm, grad_stats, layer_stats = mon.log_gradients(None, detailed=True)
    want_t = sum(p.grad.norm().item() ** 2 for (n, p), h in zip(named, has_grad) if h) ** 0.5
    want_p = sum(p.norm().item() ** 2 for (n, p), h in zip(named, has_grad) if h) ** 0.5
    assert abs(total - want_t) < 1e-4 * want_t and abs(pnorm - want_p) < 1e-4 * want_p
    assert mon.grad_norms == [total] and mon.param_norms == [pnorm]
    assert set(layer_stats) == {".".join(n.split(".")[:2]) for (n, _), h in zip(named, has_grad) if h}
    assert all(st["count"] == 1 for st in layer_stats.values())
    big = named[3][0]
    assert set(grad_stats) == {big} | {n for (n, p), h in zip(named, has_grad) if h and p.grad.norm().item() > 10.0}
    gs = grad_stats[big]
    assert abs(gs["grad_max"] - named[3][1].grad.abs().max().item()) < 1e-6 and not gs["has_nan"] and not gs["has_inf"]
    assert set(mon.layer_grad_history) == set(layer_stats)
    assert mon.check_explosion(threshold=want_t * 0.5) and mon.explosion_count == 1
    assert not mon.check_explosion(threshold=want_t * 2.0)
    named[1][1].grad[0] = float("nan")                                                   # non-finite gradients are reported
    _, _, grad_stats, _ = mon.log_gradients(None)
    assert grad_stats[named[1][0]]["has_nan"]


# ------------------------------------------------------------------------------------------------ round-2 additions
def test_per_step_scalars_do_not_leak_across_steps_when_the_host_runs_ahead():
    """lr / beta1 / bias corrections travel through a ring of pinned slots + an H2D copy in stream order, so a
    host that enqueues many graph replays ahead of the GPU cannot make step t read the scalars of step t + k."""
    named = _toy_params(5)
    ref = [p.detach().clone().requires_grad_(True) for _, p in named]
    opt = Q.FusedAdamW(named, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05)
    topt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05)
    opt.zero_grad()
    g = torch.Generator().manual_seed(2)
    for (_, p), r in zip(named, ref):
        gr = torch.randn(p.shape, generator=g).cuda()
        p.grad.copy_(gr)
        r.grad = gr.clone()
    opt.set_grad_mask([True] * len(named))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()                      # warm-up outside capture (advances the counter once, mirrored below)
    torch.cuda.current_stream().wait_stream(s)
    topt.step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    lrs = [1e-3 * (1 + 7 * (i % 3)) for i in range(24)]
    b1s = [0.85 + 0.01 * (i % 5) for i in range(24)]
    busy = torch.randn(8192, 8192, device="cuda")
    for _ in range(20):                 # ~100 ms of queued work: the host now runs far ahead of the GPU
        busy = busy @ busy * 1e-4
    for lr, b1 in zip(lrs, b1s):
        opt.param_groups[0]["lr"] = lr
        opt.param_groups[0]["betas"] = (b1, 0.999)
        opt.push_hyper()
        graph.replay()
    torch.cuda.synchronize()
    for lr, b1 in zip(lrs, b1s):
        topt.param_groups[0]["lr"] = lr
        topt.param_groups[0]["betas"] = (b1, 0.999)
        topt.step()
    # 25 chained steps of fp32 AdamW against torch's: rounding differences stay < 1e-5; ONE step taken with a neighbouring
    # step's lr (they differ by up to 7e-3) would move every element by ~1e-3
    for (n, p), r in zip(named, ref):
        assert torch.allclose(p.detach(), r.detach(), rtol=1e-4, atol=2e-5), (n, (p.detach() - r.detach()).abs().max().item())


def test_optimizer_state_dict_resumes_like_torch_adamw():
    named = _toy_params(6)
    ref = [p.detach().clone().requires_grad_(True) for _, p in named]
    opt = Q.FusedAdamW(named, lr=2e-3, betas=(0.9, 0.999), weight_decay=0.05)
    g = torch.Generator().manual_seed(4)

    def grads():
        return [torch.randn(p.shape, generator=g).cuda() for _, p in named]

    opt.set_grad_mask([True] * len(named))
    for _ in range(2):
        opt.zero_grad()
        for (_, p), gr in zip(named, grads()):
            p.grad.copy_(gr)
        opt.step()
    sd = opt.state_dict()
    # resume in torch.optim.AdamW from OUR checkpoint ...
    cur = [p.detach().clone().requires_grad_(True) for _, p in named]
    topt = torch.optim.AdamW(cur, lr=1.0)
    topt.load_state_dict(sd)
    # ... and in a fresh FusedAdamW from a torch-format checkpoint
    named2 = [(n, torch.nn.Parameter(p.detach().clone())) for n, p in named]
    opt2 = Q.FusedAdamW(named2, lr=1.0)
    opt2.load_state_dict(topt.state_dict())
    opt2.set_grad_mask([True] * len(named))
    for _ in range(2):
        gs = grads()
        opt.zero_grad()
        opt2.zero_grad()
        for (_, p), (_, p2), r, gr in zip(named, named2, cur, gs):
            p.grad.copy_(gr)
            p2.grad.copy_(gr)
            r.grad = gr.clone()
        opt.step()
        opt2.step()
        topt.step()
    for (n, p), (_, p2), r in zip(named, named2, cur):
        assert torch.allclose(p.detach(), r.detach(), rtol=2e-6, atol=1e-7), n
        assert torch.equal(p.detach(), p2.detach()), n


def test_cross_entropy_flags_out_of_range_targets_and_is_bitwise_reproducible():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(4736, 100, generator=g).cuda()
    y = torch.randint(0, 100, (4736,), generator=g).cuda()
    a = Q.cross_entropy(logits, y, label_smoothing=0.1)
    for _ in range(5):
        assert torch.equal(Q.cross_entropy(logits, y, label_smoothing=0.1), a)        # fixed-order reduction
    assert abs(a.item() - torch.nn.functional.cross_entropy(logits, y, label_smoothing=0.1).item()) < 1e-5
    Q.functional.check_labels()                                                      # all in range: no error
    bad = y.clone()
    bad[17] = 100
    Q.cross_entropy(logits, bad, label_smoothing=0.1)
    with pytest.raises(RuntimeError, match="out of range"):
        Q.functional.check_labels()
    Q.functional.check_labels()                                                      # flag was cleared
    # lam as a device scalar == lam as a float
    yb = torch.randint(0, 100, (4736,), generator=g).cuda()
    lam_t = torch.tensor(0.3, device="cuda")
    assert torch.equal(Q.cross_entropy(logits, y, 0.1, target_b=yb, lam=lam_t), Q.cross_entropy(logits, y, 0.1, target_b=yb, lam=0.3))
    lg = logits.clone().requires_grad_(True)
    (Q.cross_entropy(lg, y, label_smoothing=0.1) * 2.5).backward()
    lt = logits.clone().requires_grad_(True)
    (torch.nn.functional.cross_entropy(lt, y, label_smoothing=0.1) * 2.5).backward()
    assert torch.allclose(lg.grad, lt.grad, rtol=1e-4, atol=1e-9)


def test_graphed_step_supports_the_two_target_mixup_loss():
    import copy
    model, ocfg, sd, _ = build_model("qavitv2_c100", precision="fp32")
    model.train()
    B = 4
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, 3, 32, 32, generator=g).cuda()
    y = torch.randint(0, 100, (B,), generator=g).cuda()
    yb = torch.randint(0, 100, (B,), generator=g).cuda()
    opt = Q.FusedAdamW(model.named_parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.05, max_grad_norm=0.5)
    step = Q.GraphedTrainStep(model, opt, x, y, label_smoothing=0.1, autocast_bf16=False, warmup=2, mix=True)
    for lam, second in ((0.3, yb), (1.0, None), (0.75, yb)):
        snap = copy.deepcopy(model)
        want = Q.cross_entropy(snap(x), y, label_smoothing=0.1, target_b=second, lam=lam).item()
        got = step(x, y, y_b=second, lam=lam).item()
        assert abs(got - want) < 2e-5 * max(1.0, abs(want)), (lam, got, want)


def test_normalize_batch_is_bit_identical_to_totensor_plus_normalize():
    g = torch.Generator().manual_seed(0)
    raw = torch.randint(0, 256, (37, 32, 32, 3), generator=g, dtype=torch.uint8)          # dataset layout [B, H, W, C]
    mean, std = Q.evalutil.CIFAR100_MEAN, Q.evalutil.CIFAR100_STD
    t = raw.permute(0, 3, 1, 2).float().div(255)                                          # transforms.ToTensor
    want = (t - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)  # transforms.Normalize: sub_ then div_
    got = Q.normalize_batch(raw.cuda(), mean, std)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(Q.normalize_batch(raw.cuda(), mean, std, hflip=True).cpu(), want.flip(-1))      # RandomHorizontalFlip(p=1)
    assert torch.equal(Q.normalize_batch(raw.permute(0, 3, 1, 2).contiguous().cuda(), mean, std).cpu(), want)
    assert torch.equal(Q.normalize_batch(t.cuda(), mean, std).cpu(), want)
    v = Q.tta_views(raw.cuda())
    assert len(v) == 2 and torch.equal(v[1], v[0].flip(-1))


def test_validate_tta_matches_the_reference_algorithm():
    """HQAViT_C100_Finetune.py:345-384 restated with host-side accumulation, against the device-resident drop-in."""
    model, ocfg, sd, _ = build_model("hqavit_c100", precision="fp32")
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(6, 3, 32, 32, generator=g) for _ in range(3)]
    ys = [torch.randint(0, 100, (6,), generator=g) for _ in range(3)]
    loaders = [list(zip(xs, ys)), [(x.flip(-1), y) for x, y in zip(xs, ys)], [(x + 0.05, y) for x, y in zip(xs, ys)]]
    model.eval()
    with torch.no_grad():
        preds = [torch.cat([torch.softmax(model(x.cuda()), 1).cpu() for x, _ in ld]) for ld in loaders]
        # make the check non-trivial: targets = the ensemble's own prediction on half of the samples
        ens = torch.stack(preds).mean(0).argmax(1)
    tgt = torch.cat(ys)
    tgt[::2] = ens[::2]
    loaders = [[(x, tgt[6 * i:6 * i + 6]) for i, (x, _) in enumerate(ld)] for ld in loaders]
    want = 100.0 * ens.eq(tgt).sum().item() / tgt.numel()
    got = Q.validate_tta(model, loaders)
    assert abs(got - want) < 1e-9 and want >= 50.0
    assert not model.training
    assert Q.validate_tta(model, []) == 0.0


def test_forward_hook_on_patch_embed_proj_fires_with_the_conv_activation():
    """Grad-CAM of the reference (test_hqa.py:241-275) hooks model.patch_embed.proj: the hook must see the [B, d, H/p, W/p]
    activation and its gradient, and the hooked forward must give the same logits as the fused one."""
    model, ocfg, sd, _ = build_model("hqavit_c100", precision="fp32")
    model.eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 3, 32, 32, generator=g).cuda()
    with torch.no_grad():
        plain = model(x)
    seen = {}

    def hook(module, inp, out):
        seen["act"] = out
        out.register_hook(lambda gr: seen.__setitem__("grad", gr))

    h = model.patch_embed.proj.register_forward_hook(hook)
    out = model(x)
    out[:, 3].sum().backward()
    h.remove()
    assert torch.allclose(out, plain, rtol=1e-3, atol=1e-4), (out - plain).abs().max().item()   # separate GEMM + LayerNorm kernels vs the fused one
    conv = torch.nn.functional.conv2d(x, model.patch_embed.proj.weight, model.patch_embed.proj.bias, stride=4)
    assert seen["act"].shape == (2, 192, 8, 8) and torch.allclose(seen["act"], conv, rtol=1e-4, atol=1e-5)
    assert seen["grad"].shape == (2, 192, 8, 8) and seen["grad"].abs().sum().item() > 0
    with torch.no_grad():
        assert torch.equal(model(x), plain)          # hook removed: back on the fused path


def test_match_autocast_output_dtype_is_opt_in():
    model, ocfg, sd, _ = build_model("qavitv2_c100", precision="auto")
    model.eval()
    x = torch.randn(2, 3, 32, 32).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a = model(x)
        model.match_autocast_output_dtype = True
        b = model(x)
    assert a.dtype == torch.float32 and b.dtype == torch.bfloat16       # the reference's head Linear returns bf16 under autocast
    assert torch.equal(a.to(torch.bfloat16), b)


def test_graphed_step_prefetch_pipeline_equals_direct_copies():
    """GraphedTrainStep.prefetch(): the next batch's pinned-host -> device copy runs on a copy stream under the current step; the
    step sequence must see exactly the batches it would see with the copy in front of each step."""
    import copy
    base, ocfg, sd, _ = build_model("qavitv2_c100", precision="fp32")
    base.train()
    B, n = 4, 4
    g = torch.Generator().manual_seed(21)
    xs = [torch.randn(B, 3, 32, 32, generator=g).pin_memory() for _ in range(n)]
    ys = [torch.randint(0, 100, (B,), generator=g).pin_memory() for _ in range(n)]

    def make():
        m = copy.deepcopy(base)
        opt = Q.FusedAdamW(m.named_parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, max_grad_norm=0.5)
        return Q.GraphedTrainStep(m, opt, xs[0].cuda(), ys[0].cuda(), label_smoothing=0.1, autocast_bf16=False, warmup=2)
    a, b = make(), make()
    direct = [a(xs[i], ys[i]).item() for i in range(n)]
    piped = []
    b.prefetch(xs[0], ys[0])
    for i in range(n):
        loss = b()
        if i + 1 < n:
            b.prefetch(xs[i + 1], ys[i + 1])
        piped.append(loss.item())
    assert len(set(round(v, 3) for v in direct)) > 1          # the batches differ, so the losses do
    for d, p in zip(direct, piped):
        assert abs(d - p) < 1e-4 * max(1.0, abs(d)), (direct, piped)
