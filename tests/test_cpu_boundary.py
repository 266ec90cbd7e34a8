"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol of
include/qavit_b200.h, the host-side workspace/layout logic runs, the module trees carry exactly the reference's
state_dict schema, and the product path refuses CPU tensors instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

import qavit_b200 as Q
from oracle import qavit_oracle as O
from qavit_b200 import _lib
from util import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qavit_b200.h")).read()
    declared = set(re.findall(r"\b(qavit_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("qavit_block_cfg")
    assert declared, "no declarations parsed"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/qavit_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.qavit_abi_version() == 3


def test_param_table_names_exist_in_reference_schema():
    sch = O.state_schema(O.OracleConfig(family="hqavit"))
    assert _lib.QP_COUNT == 91
    for name, scope in _lib.PARAMS:
        key = {0: "stage1_blocks.0.quad_block." + name, 1: "stage1_blocks.0." + name, 2: "global_bank." + name}[scope]
        if name == "ccf_ffn.dwconv.dwconv.bias":
            continue                      # only in QAViT.py / QAViTv2.py
        assert key in sch, key


def test_workspace_query_runs_on_host():
    c = _lib.BlockCfg()
    c.batch, c.tokens, c.tokens_full, c.token_learner = 256, 16, 64, 1
    c.dim, c.heads, c.bank_size, c.groups, c.window, c.linformer_k, c.msda_seq_len = 192, 4, 16, 6, 4, 32, 128
    c.n_dilations, c.pool_stride = 2, 2
    c.dilations[0], c.dilations[1] = 1, 2
    c.compress_dim, c.bottleneck_hidden, c.ffn_hidden = 48, 96, 96
    c.train, c.dtype = 1, 1
    a, b = C.c_size_t(0), C.c_size_t(0)
    assert _lib.lib.qavit_block_workspace(C.byref(c), C.byref(a), C.byref(b)) == 0
    assert a.value > 256 * 16 * 192 * 2 * 10 and b.value > 0
    c.window = 3                          # 4x4 grid is not divisible by 3: the reference cannot run this either (H:466)
    assert _lib.lib.qavit_block_workspace(C.byref(c), C.byref(a), C.byref(b)) != 0
    assert b"window" in _lib.lib.qavit_last_error()


@pytest.mark.parametrize("case", list(CASES))
def test_state_dict_schema_matches_reference(case):
    from util import build_model
    model, ocfg, sd, _ = build_model(case, device="cpu")
    want = O.with_bank_aliases(sd, ocfg)
    have = model.state_dict()
    assert set(have) == set(want)
    for k, v in have.items():
        assert tuple(v.shape) == tuple(want[k].shape), k
    # aliased bank entries share storage with the top-level bank (one GlobalTokenBank under every branch)
    P = O.block_prefixes(ocfg)[-1]
    assert have[f"{P}.cga.global_bank.global_k"].data_ptr() == have["global_bank.global_k"].data_ptr()
    n_named = len(dict(model.named_parameters()))
    assert n_named == len(O.trainable_keys(ocfg))


def test_no_cpu_fallback():
    from util import build_model
    model, ocfg, _, _ = build_model("qavitv2_c100", device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        Q.cross_entropy(torch.zeros(2, 10), torch.zeros(2, dtype=torch.long))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "qa-vit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} references the oracle"


def test_transfer_helpers_follow_the_reference_recipes():
    """adjust_positional_embedding (HQAViT_Tiny_stl10.py:250-282) and the head-swap load (HQAViT_Tiny_Cifar10.py:445-453)
    on the drop-in module tree (host-side tensor surgery, no kernel involved)."""
    import torch.nn.functional as F
    src = Q.HQAViT(Q.HQAViTConfig(dropout=0.0, drop_path=0.0))
    dst = Q.HQAViT(Q.HQAViTConfig(num_classes=10, dropout=0.0, drop_path=0.0))
    loaded, skipped = Q.load_pretrained_except_head(dst, src.state_dict())
    assert skipped == 2 and loaded == len(src.state_dict()) - 2
    assert torch.equal(dst.stage3_blocks[0].quad_block.swa.qkv.weight, src.stage3_blocks[0].quad_block.swa.qkv.weight)
    assert dst.head.weight.shape == (10, 192)
    old = dst.pos_embed.detach().clone()
    assert not Q.adjust_positional_embedding(dst, 32)
    assert Q.adjust_positional_embedding(dst, 96)
    assert dst.pos_embed.shape == (1, 576, 192) and isinstance(dst.pos_embed, torch.nn.Parameter)
    want = F.interpolate(old.reshape(1, 8, 8, 192).permute(0, 3, 1, 2), size=(24, 24), mode="bicubic", align_corners=False)
    assert torch.equal(dst.pos_embed.detach(), want.permute(0, 2, 3, 1).reshape(1, 576, 192))
    ref_path = "/root/reference"
    if os.path.isdir(ref_path):        # same result as the reference's own function applied to our module
        import importlib.util, sys, types
        for stub in ("matplotlib", "matplotlib.pyplot", "seaborn"):
            sys.modules.setdefault(stub, types.ModuleType(stub))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].use = lambda *a, **k: None
        try:
            spec = importlib.util.spec_from_file_location("ref_stl10", os.path.join(ref_path, "HQAViT_Tiny_stl10.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.path.insert(0, ref_path)
            spec.loader.exec_module(mod)
        except Exception as e:          # optional dependency of that script missing in this image
            pytest.skip(f"reference script not importable here: {e}")
        other = Q.HQAViT(Q.HQAViTConfig(num_classes=10, dropout=0.0, drop_path=0.0))
        other.pos_embed.data.copy_(old)
        mod.adjust_positional_embedding(other, 96)
        assert torch.equal(other.pos_embed.detach(), dst.pos_embed.detach())
