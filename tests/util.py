"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from oracle import qavit_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def golden(case):
    return np.load(os.path.join(HERE, "golden", case + ".npz"))


def inputs(ocfg, B, seed=1234):
    """Same generator recipe as tests/golden/make_golden.py::inputs."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, ocfg.in_channels, ocfg.img_size, ocfg.img_size, generator=g)
    y = torch.randint(0, ocfg.num_classes, (B,), generator=g)
    return x, y


CASES = {
    # name: (OracleConfig kwargs, model family, ctor kwargs, batch used for the golden vectors)
    "hqavit_c100": (dict(family="hqavit"), "hqavit", {}, 4),
    "qavitv2_c100": (dict(family="qavit_v2"), "qavit", dict(variant="v2"), 2),
    "qavit_v1_c10": (dict(family="qavit_v1", num_classes=10, dwconv_bias=True), "qavit", dict(variant="v1"), 2),
    "qavit_v1_224": (dict(family="qavit_v1", img_size=224, patch_size=16, window_size=7, dilation_factors=(1, 2, 3), linformer_k=64,
                          dwconv_bias=True), "qavit", dict(variant="v1"), 2),
    "qavitv2b_224": (dict(family="qavit_v2", img_size=224, patch_size=16, window_size=7, dilation_factors=(1, 2, 3), linformer_k=64,
                          dwconv_bias=True), "qavit", dict(variant="v2b"), 2),
    "hqavit_stl96": (dict(family="hqavit", img_size=96, built_img_size=32, num_classes=10), "hqavit", {}, 2),
    "hqavitv2_c100": (dict(family="hqavit", stem="v2"), "hqavit", dict(variant="v2"), 2),
    "hqavit_tinyin": (dict(family="hqavit", img_size=64, num_classes=200, depth=12, num_learned_tokens=64,
                           stage_depths=(2, 2, 6, 2)), "hqavit", dict(square_tokens=True), 2),
}


def build_model(case, device="cuda", precision="fp32"):
    """Our drop-in module for `case`, loaded with the synthetic weights, dropout disabled."""
    import qavit_b200 as Q
    okw, fam, ckw, B = CASES[case]
    ocfg = O.OracleConfig(**okw)
    common = dict(img_size=ocfg.built_img_size or ocfg.img_size, patch_size=ocfg.patch_size, num_classes=ocfg.num_classes, depth=ocfg.depth,
                  dropout=0.0, drop_path=0.0, window_size=ocfg.window_size, dilation_factors=tuple(ocfg.dilation_factors),
                  linformer_k=ocfg.linformer_k)
    if fam == "hqavit":
        cfg = Q.HQAViTConfig(num_learned_tokens=ocfg.num_learned_tokens, **common)
        model = Q.HQAViT(cfg, stage_depths=ocfg.stage_depths, **ckw)
        for n in ("fuse2", "fuse3", "fuse4"):
            getattr(model, n).cat_mlp[3].p = 0.0
        if model.variant == "v2":                      # hard-coded DropPath(0.1) of the v2 stem's blocks
            for blk in model.cnn_stem.blocks_in_order():
                blk.drop_path_rate = 0.0
    else:
        cfg = Q.QAViTConfig(**common)
        model = Q.QAViT(cfg, **ckw)
    sd = O.synthetic_state(ocfg)
    if ocfg.built_img_size and ocfg.built_img_size != ocfg.img_size:      # the STL-10 transfer recipe through OUR helpers
        Q.adjust_positional_embedding(model, ocfg.img_size)
    model.load_state_dict(O.with_bank_aliases(sd, ocfg), strict=True)
    model = model.to(device)
    model.set_precision(precision)
    return model, ocfg, sd, B


def rel_l2(a, b, floor=0.0):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return ((a - b).norm() / (b.norm() + floor)).item()


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
