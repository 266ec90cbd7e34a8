"""Host-side logic of FusedAdamW that needs no GPU: torch.optim.AdamW-compatible state_dict round trips (the reference
checkpoints optimizer.state_dict(), HQAViT_CIFAR100.py:1692 / 1728) and the name-derived default of the has-gradient mask."""
import torch

import qavit_b200 as Q


def _toy():
    torch.manual_seed(0)
    m = torch.nn.ModuleDict({"a": torch.nn.Linear(5, 3), "swa": torch.nn.ModuleDict({"norm": torch.nn.LayerNorm(3)}),
                             "b": torch.nn.Linear(3, 2, bias=False)})
    return m


def test_state_dict_loads_torch_adamw_state_and_round_trips():
    m = _toy()
    ref = torch.optim.AdamW(m.parameters(), lr=3e-4, betas=(0.95, 0.999), weight_decay=0.06)
    for _ in range(3):
        for p in m.parameters():
            p.grad = torch.randn_like(p)
        ref.step()
    sd = ref.state_dict()

    m2 = _toy()
    opt = Q.FusedAdamW(m2.named_parameters(), lr=1.0, betas=(0.9, 0.9), weight_decay=0.0)   # CPU: host logic only
    opt.load_state_dict(sd)
    assert opt.step_count == 3
    g = opt.param_groups[0]
    assert g["lr"] == 3e-4 and tuple(g["betas"]) == (0.95, 0.999) and g["weight_decay"] == 0.06
    offs = opt.seg_off.tolist()
    for i, p in enumerate(m2.parameters()):
        assert torch.equal(opt.exp_avg[offs[i]:offs[i] + p.numel()].view(p.shape), sd["state"][i]["exp_avg"])
        assert torch.equal(opt.exp_avg_sq[offs[i]:offs[i] + p.numel()].view(p.shape), sd["state"][i]["exp_avg_sq"])

    # export: every parameter with a gradient flag carries {step, exp_avg, exp_avg_sq}; torch.optim.AdamW accepts it
    opt._sync_flags([True] * len(opt.names))
    out = opt.state_dict()
    assert set(out["state"]) == set(range(len(opt.names)))
    ref2 = torch.optim.AdamW(_toy().parameters(), lr=1.0)
    ref2.load_state_dict(out)
    for i in range(len(opt.names)):
        assert torch.equal(ref2.state_dict()["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert float(ref2.state_dict()["state"][i]["step"]) == 3.0
    assert ref2.state_dict()["param_groups"][0]["lr"] == 3e-4

    # and back into a fresh FusedAdamW
    opt3 = Q.FusedAdamW(_toy().named_parameters(), lr=1.0)
    opt3.load_state_dict(out)
    assert opt3.step_count == 3 and torch.equal(opt3.exp_avg, opt.exp_avg) and torch.equal(opt3.exp_avg_sq, opt.exp_avg_sq)


def test_state_dict_skips_parameters_without_gradient_like_torch():
    m = _toy()
    opt = Q.FusedAdamW(m.named_parameters(), lr=1e-3)
    opt.step_count = 2
    opt._sync_flags([n != "swa.norm.weight" for n in opt.names])
    sd = opt.state_dict()
    assert opt.names.index("swa.norm.weight") not in sd["state"]
    assert len(sd["state"]) == len(opt.names) - 1


def test_attached_gradients_default_to_the_reference_trainable_set():
    """attach_grads() makes every .grad non-None; without an explicit mask the parameters the reference never trains
    (bank write_*, branch .norm: grad stays None there, AdamW skips them incl. weight decay) must stay flagged off."""
    m = _toy()
    opt = Q.FusedAdamW(m.named_parameters(), lr=1e-3)
    opt.zero_grad()
    opt._sync_flags()
    flags = dict(zip(opt.names, (int(f) & 1 for f in opt._flags_host.tolist())))
    assert flags["swa.norm.weight"] == 0 and flags["swa.norm.bias"] == 0
    assert flags["a.weight"] == 1 and flags["b.weight"] == 1
    # a real gradient tensor (not the attached view) always counts
    m["swa"]["norm"].weight.grad = torch.ones(3)
    opt._sync_flags()
    assert int(opt._flags_host[opt.names.index("swa.norm.weight")]) & 1 == 1


def test_cpu_optimizer_refuses_to_step():
    opt = Q.FusedAdamW(_toy().named_parameters(), lr=1e-3)
    import pytest
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.clip()
