"""Import the UNMODIFIED reference model scripts (test / bench infrastructure -- never imported by qa-vit_b200/).

Search order: baseline/_ref/ (the copy `python baseline/install_ref.py` makes; it travels to the GPU box), then
/root/reference (build container).  The reference's scripts import matplotlib at top level in two files; a stub is
injected when it is absent.  Nothing in the reference is edited: the flash-attn switch is the module's own global.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.path.join(HERE, "_ref"), os.environ.get("QAVIT_REFERENCE", "/root/reference")]

# reference module -> (model class, config class)
MODELS = {
    "HQAViT_CIFAR100": ("HQAViT", "HQAViTConfig"),
    "HQAViTv2_CIFAR100": ("HQAViT", "HQAViTConfig"),
    "HQAViT_IN_Tiny": ("HQAViT", "HQAViTConfig"),
    "QAViT": ("QAViT", "QAViTConfig"),
    "QAViTv2": ("QAViT", "QAViTConfig"),
    "QAViTv2_CIFAR100": ("QAViT", "QAViTConfig"),
    "QAViTV2_EXTREME": ("QAViT", "QAViTConfig"),
}


def ref_dir():
    for d in _CANDIDATES:
        if d and os.path.isfile(os.path.join(d, "HQAViT_CIFAR100.py")):
            return d
    return None


def available() -> bool:
    return ref_dir() is not None


def import_reference(mod_name: str, flash=False):
    """The reference module `mod_name`; flash=False pins efficient_attention to its SDPA branch (H:359-392), flash=None
    leaves the module's own HAS_FLASH_ATTN probe alone."""
    d = ref_dir()
    if d is None:
        raise ImportError("live reference not available: run `python baseline/install_ref.py` in the build container")
    sys.dont_write_bytecode = True
    if d not in sys.path:
        sys.path.insert(0, d)
    for stub in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if stub not in sys.modules:
            try:
                __import__(stub)
            except Exception:
                m = types.ModuleType(stub)
                m.use = lambda *a, **k: None
                sys.modules[stub] = m
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mod = __import__(mod_name)
    if flash is not None and hasattr(mod, "HAS_FLASH_ATTN"):
        mod.HAS_FLASH_ATTN = bool(flash) and mod.HAS_FLASH_ATTN
    return mod


def build(mod_name: str, seed=42, deterministic=True, **cfg_over):
    """(module, model): the reference model built under torch.manual_seed(seed) (the reference's own seed, H:1770).
    deterministic: dropout = drop_path = 0 and SplitFusion's hard-coded Dropout(0.1) (H:930) silenced."""
    import torch
    mod = import_reference(mod_name)
    cls, cfg_cls = MODELS[mod_name]
    if deterministic:
        cfg_over = dict(dropout=0.0, drop_path=0.0, **cfg_over)
    torch.manual_seed(seed)
    model = getattr(mod, cls)(getattr(mod, cfg_cls)(**cfg_over))
    if deterministic:
        for n in ("fuse2", "fuse3", "fuse4"):
            if hasattr(model, n):
                getattr(model, n).cat_mlp[3].p = 0.0
    return mod, model
