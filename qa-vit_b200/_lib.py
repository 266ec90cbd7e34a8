"""ctypes binding of libqavit_b200.so (the C ABI declared in include/qavit_b200.h).

There is no CPU fallback: if the library is missing, importing this module raises; if a call fails the
C side's message is raised as RuntimeError (the reference surfaces failures as Python RuntimeError too,
e.g. the OOM handling at QAViTv2.py:1196-1202)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QAVIT_LIB") or os.path.join(_HERE, "libqavit_b200.so")   # QAVIT_LIB: A/B builds of the same ABI (tools/)


class BlockCfg(C.Structure):
    """Mirror of ``qavit_block_cfg`` (include/qavit_b200.h)."""
    _fields_ = [
        ("batch", C.c_int32), ("tokens", C.c_int32), ("tokens_full", C.c_int32), ("token_learner", C.c_int32),
        ("dim", C.c_int32), ("heads", C.c_int32), ("bank_size", C.c_int32), ("groups", C.c_int32),
        ("window", C.c_int32), ("linformer_k", C.c_int32), ("msda_seq_len", C.c_int32),
        ("n_dilations", C.c_int32), ("dilations", C.c_int32 * 4), ("pool_stride", C.c_int32),
        ("compress_dim", C.c_int32), ("bottleneck_hidden", C.c_int32), ("ffn_hidden", C.c_int32),
        ("ffn_v1", C.c_int32), ("dwconv_bias", C.c_int32), ("bank_v1", C.c_int32), ("train", C.c_int32),
        ("dtype", C.c_int32), ("dropout", C.c_float), ("drop_path", C.c_float), ("tokens_out", C.c_int32),
    ]


class LateralCfg(C.Structure):
    """Mirror of ``qavit_lateral_cfg``."""
    _fields_ = [
        ("batch", C.c_int32), ("img_size", C.c_int32), ("in_channels", C.c_int32),
        ("c_stem", C.c_int32), ("c2", C.c_int32), ("c3", C.c_int32), ("c4", C.c_int32),
        ("rrcv_channels", C.c_int32), ("rrcv_blocks", C.c_int32), ("dim", C.c_int32), ("grid", C.c_int32),
        ("train", C.c_int32), ("dtype", C.c_int32), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
        ("stem_kind", C.c_int32), ("stem_drop_path", C.c_float * 7), ("rng", C.c_void_p),
    ]


class SplitFusionCfg(C.Structure):
    """Mirror of ``qavit_splitfusion_cfg``."""
    _fields_ = [("rows", C.c_longlong), ("dim", C.c_int32), ("dtype", C.c_int32), ("train", C.c_int32), ("drop_p", C.c_float)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()' or make -C qa-vit_b200/csrc).  qavit_b200 has no CPU / eager fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_SIGS = {
    "qavit_last_error": (C.c_char_p, []),
    "qavit_abi_version": (_i, []),
    "qavit_launch_count": (_ll, []),
    "qavit_block_param_name": (C.c_char_p, [_i, C.POINTER(_i)]),
    "qavit_block_workspace": (_i, [C.POINTER(BlockCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "qavit_block_forward": (_i, [C.POINTER(BlockCfg), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_block_backward": (_i, [C.POINTER(BlockCfg), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_patch_embed_scratch_bytes": (C.c_size_t, [_i, _i, _i, _i, _i]),
    "qavit_patch_embed_forward": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "qavit_patch_embed_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "qavit_head_forward": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "qavit_head_backward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_dropout_forward": (_i, [_vp, _vp, _ll, _f, _vp, _vp, _vp]),
    "qavit_dropout_backward": (_i, [_vp, _vp, _ll, _f, _vp, _vp]),
    "qavit_cross_entropy": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "qavit_scale_by_scalar": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "qavit_memset_zero": (_i, [_vp, C.c_size_t, _vp]),
    "qavit_clip_grads": (_i, [_vp, _vp, _vp, _i, _f, _f, _vp, _ll, _vp]),
    "qavit_clip_grads_scaled": (_i, [_vp, _vp, _vp, _i, _f, _f, _f, _vp, _ll, _vp]),
    "qavit_normalize_images": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "qavit_scaled_copy": (_i, [_vp, _f, _ll, _vp, _vp]),
    "qavit_adamw_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _ll, _vp]),
    "qavit_adamw_ema_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _ll, _vp]),
    "qavit_segment_norms": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "qavit_batch_mix": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp]),
    "qavit_layer_norm_forward": (_i, [_vp, _i, _ll, _i, _vp, _vp, _f, _vp, _vp, _vp]),
    "qavit_layer_norm_backward": (_i, [_vp, _i, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_dwconv_forward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "qavit_dwconv_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "qavit_linear_forward": (_i, [_vp, _i, _ll, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "qavit_linear_backward": (_i, [_vp, _vp, _i, _ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_lateral_param_count": (_i, [C.POINTER(LateralCfg)]),
    "qavit_lateral_param_name": (C.c_char_p, [C.POINTER(LateralCfg), _i]),
    "qavit_lateral_workspace": (_i, [C.POINTER(LateralCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "qavit_lateral_forward": (_i, [C.POINTER(LateralCfg), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_lateral_backward": (_i, [C.POINTER(LateralCfg), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_lateral_forward_parts": (_i, [C.POINTER(LateralCfg), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint]),
    "qavit_lateral_backward_parts": (_i, [C.POINTER(LateralCfg), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                          C.c_uint]),
    "qavit_dp_available": (_i, []),
    "qavit_dp_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "qavit_dp_allreduce_sum": (_i, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t), _i, _vp]),
    "qavit_splitfusion_param_name": (C.c_char_p, [_i]),
    "qavit_splitfusion_workspace": (_i, [C.POINTER(SplitFusionCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "qavit_splitfusion_forward": (_i, [C.POINTER(SplitFusionCfg), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_splitfusion_backward": (_i, [C.POINTER(SplitFusionCfg), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qavit_test_gemm_nt": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "qavit_test_gemm_epi": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "qavit_test_gemm_tn": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "qavit_test_tokens_fused": (_i, [_i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "qavit_test_cmp_fused": (_i, [_i, _ll, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "qavit_test_ffn_mid": (_i, [_i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "qavit_convert_weight": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
}
EXPORTS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == the .so does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int) -> None:
    if status != 0:
        raise RuntimeError("qavit_b200: " + lib.qavit_last_error().decode("utf-8", "replace"))


def param_table():
    """[(state_dict suffix, scope)] in QP_* order; scope 0 = quad block, 1 = TokenLearner wrapper, 2 = bank."""
    out = []
    i = 0
    while True:
        scope = _i(0)
        name = lib.qavit_block_param_name(i, C.byref(scope))
        if name is None:
            break
        out.append((name.decode(), scope.value))
        i += 1
    return out


PARAMS = param_table()
QP_COUNT = len(PARAMS)
QP = {name if scope != 2 else "bank." + name: i for i, (name, scope) in enumerate(PARAMS)}


def lateral_param_names(cfg: LateralCfg):
    """state_dict names (relative to the HQAViT module) of the lateral path's tensors, in the C side's order."""
    n = lib.qavit_lateral_param_count(C.byref(cfg))
    if n < 0:
        check(1)
    return [lib.qavit_lateral_param_name(C.byref(cfg), i).decode() for i in range(n)]


SPLITFUSION_PARAMS = []
while lib.qavit_splitfusion_param_name(len(SPLITFUSION_PARAMS)) is not None:
    SPLITFUSION_PARAMS.append(lib.qavit_splitfusion_param_name(len(SPLITFUSION_PARAMS)).decode())
