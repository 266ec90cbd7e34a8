"""Generate tests/golden/*.npz from the LIVE reference modules (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference, loads the deterministic synthetic
weights of oracle.qavit_oracle.synthetic_state() into the reference nn.Modules
(strict=True, including the aliased bank keys), runs them on seeded inputs and stores
compact outputs: eval logits, train-mode logits / loss / per-parameter gradient
fingerprints (L2 norm, sum, first 8 values) / post-forward bank state, and the state after
3 clip+AdamW steps.  The .npz files are committed; the GPU box never reads /root/reference.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("QAVIT_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True

from oracle import qavit_oracle as O  # noqa: E402

CASES = {
    # name: (reference module, model class, config class, config overrides, OracleConfig kwargs, batch)
    "hqavit_c100": ("HQAViT_CIFAR100", "HQAViT", "HQAViTConfig", {}, dict(family="hqavit"), 4),
    "qavitv2_c100": ("QAViTv2_CIFAR100", "QAViT", "QAViTConfig", {}, dict(family="qavit_v2"), 2),
    "qavit_v1_c10": ("QAViT", "QAViT", "QAViTConfig",
                     dict(img_size=32, patch_size=4, num_classes=10, window_size=4, dilation_factors=(1, 2),
                          linformer_k=32),
                     dict(family="qavit_v1", num_classes=10, dwconv_bias=True), 2),
    # the files' OWN defaults: 224 x 224 / patch 16, 196 tokens, 7 x 7 windows (49 tokens), dilations (1, 2, 3), Linformer k = 64
    # (QAViT.py:39-56, QAViTv2.py:43-60); MSDA pools 270 multi-scale tokens to 135 and truncates to 128 (QAViT.py:413-420)
    "qavit_v1_224": ("QAViT", "QAViT", "QAViTConfig", {},
                     dict(family="qavit_v1", img_size=224, patch_size=16, window_size=7, dilation_factors=(1, 2, 3), linformer_k=64,
                          dwconv_bias=True), 2),
    "qavitv2b_224": ("QAViTv2", "QAViT", "QAViTConfig", {},
                     dict(family="qavit_v2", img_size=224, patch_size=16, window_size=7, dilation_factors=(1, 2, 3), linformer_k=64,
                          dwconv_bias=True), 2),
    # STL-10 fine-tuning recipe (HQAViT_Tiny_stl10.py:250-282, 405-412): the 32 x 32 CIFAR-100 model fed 96 x 96 images after
    # adjust_positional_embedding(model, 96) and a 10-class head: block 0's TokenLearner sees 576 tokens, TokenUpMix still emits 64,
    # the lateral path runs on 24 x 24 maps and LMFAdapter resizes them bilinearly to 8 x 8 (H:839-843)
    "hqavit_stl96": ("HQAViT_CIFAR100", "HQAViT", "HQAViTConfig", {},
                     dict(family="hqavit", img_size=96, built_img_size=32, num_classes=10), 2),
    # HQAViTv2_CIFAR100.py: the same model around a ConvNeXt-patchify stem (LayerNorm([C, 8, 8]), LayerScale, DropPath), V:753-833
    "hqavitv2_c100": ("HQAViTv2_CIFAR100", "HQAViT", "HQAViTConfig", {}, dict(family="hqavit", stem="v2"), 2),
    "hqavit_tinyin": ("HQAViT_IN_Tiny", "HQAViT", "HQAViTConfig", {},
                      dict(family="hqavit", img_size=64, num_classes=200, depth=12, num_learned_tokens=64,
                           stage_depths=(2, 2, 6, 2)), 2),
}


def import_reference(mod_name):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for stub in ("matplotlib", "matplotlib.pyplot"):       # HQAViT_IN_Tiny imports them at top level
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.use = lambda *a, **k: None
            sys.modules[stub] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    return __import__(mod_name)


def build_reference(case):
    mod_name, cls, cfg_cls, over, okw, B = CASES[case]
    mod = import_reference(mod_name)
    mod.HAS_FLASH_ATTN = False
    rcfg = getattr(mod, cfg_cls)(dropout=0.0, drop_path=0.0, **over)
    model = getattr(mod, cls)(rcfg)
    for n in ("fuse2", "fuse3", "fuse4"):                   # hidden Dropout(0.1), H:930
        if hasattr(model, n):
            getattr(model, n).cat_mlp[3].p = 0.0
    if hasattr(model, "cnn_stem"):                          # hard-coded DropPath(0.1) of HQAViTv2's stem blocks, V:787-799
        for m in model.cnn_stem.modules():
            if hasattr(m, "drop_path"):
                m.drop_path = torch.nn.Identity()
    ocfg = O.OracleConfig(**okw)
    if ocfg.built_img_size and ocfg.built_img_size != ocfg.img_size:      # the reference's own transfer recipe
        stl = import_reference("HQAViT_Tiny_stl10")
        stl.adjust_positional_embedding(model, ocfg.img_size)
        model.head = torch.nn.Linear(model.head.in_features, ocfg.num_classes)
    return model, ocfg, B


def inputs(ocfg, B, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, ocfg.in_channels, ocfg.img_size, ocfg.img_size, generator=g)
    y = torch.randint(0, ocfg.num_classes, (B,), generator=g)
    return x, y


def fingerprint(g):
    if g is None:
        return np.full(11, np.nan, dtype=np.float64)
    f = g.detach().double().flatten()
    head = torch.zeros(8, dtype=torch.float64)
    head[: min(8, f.numel())] = f[:8]
    return np.concatenate([[f.norm().item(), f.sum().item(), float(f.numel())], head.numpy()])


def run_case(case, ls=0.1, steps=3, lr=6e-4, wd=0.06):
    model, ocfg, B = build_reference(case)
    sd = O.synthetic_state(ocfg)
    missing = model.load_state_dict(O.with_bank_aliases(sd, ocfg), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x, y = inputs(ocfg, B)
    out = {"keys": np.array(O.trainable_keys(ocfg))}

    model.eval()
    with torch.no_grad():
        out["eval_logits"] = model(x).numpy()

    model.train()
    model.zero_grad(set_to_none=True)
    logits = model(x)
    loss = torch.nn.functional.cross_entropy(logits, y, label_smoothing=ls)
    loss.backward()
    named = dict(model.named_parameters())
    out["train_logits"] = logits.detach().numpy()
    out["train_loss"] = np.array(loss.item())
    out["grad_fp"] = np.stack([fingerprint(named[k].grad) for k in O.trainable_keys(ocfg)])
    out["bank_k_after_fwd"] = model.global_bank.global_k.detach().numpy().copy()
    out["bank_v_after_fwd"] = model.global_bank.global_v.detach().numpy().copy()
    if hasattr(model.global_bank, "update_count"):
        out["update_count_after_fwd"] = np.array(int(model.global_bank.update_count))

    # three optimizer steps from a fresh copy (bank carry-over + clip + AdamW), H:1413-1439
    model.load_state_dict(O.with_bank_aliases(sd, ocfg), strict=True)
    opt = torch.optim.AdamW(model.parameters(), lr=lr, betas=(0.9, 0.999), weight_decay=wd)
    losses = []
    for s in range(steps):
        xs, ys = inputs(ocfg, B, seed=1234 + s)
        opt.zero_grad(set_to_none=True)
        l = torch.nn.functional.cross_entropy(model(xs), ys, label_smoothing=ls)
        l.backward()
        for n, p in model.named_parameters():
            if p.grad is not None and ("cnn_stem" in n or "dwconv" in n):
                torch.nn.utils.clip_grad_norm_([p], 0.1)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
        losses.append(l.item())
    out["step_losses"] = np.array(losses)
    named = dict(model.named_parameters())
    out["param_fp_after_steps"] = np.stack([fingerprint(named[k]) for k in O.trainable_keys(ocfg)])
    model.eval()
    with torch.no_grad():
        out["eval_logits_after_steps"] = model(x).numpy()
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    for case in (sys.argv[1:] or CASES):
        res = run_case(case)
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), case + ".npz")
        np.savez_compressed(path, **res)
        print(case, "->", path, os.path.getsize(path), "bytes", "loss", float(res["train_loss"]))
