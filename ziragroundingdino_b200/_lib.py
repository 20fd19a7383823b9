"""ctypes binding of libmsda_b200.so -- the C ABI declared in include/msda_b200.h.

The library is built in-tree by ``ziragroundingdino_b200/csrc/build.sh`` (``__graft_entry__.build()``).
There is no fallback of any kind: if the shared object is missing or a call fails, an exception is
raised (the reference op raises ``RuntimeError`` through ``AT_ASSERTM`` / ``AT_ERROR``,
csrc/MsDeformAttn/ms_deform_attn_cuda.cu:29-53, ms_deform_attn.h:39).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libmsda_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "msda_b200.h")

_lib = None

_vp, _i, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
_FWD = [_vp] * 5 + [_i] * 7 + [_vp, _vp]
_BWD = [_vp] * 6 + [_i] * 7 + [_vp, _vp, _vp, _i, _vp]


def declared_symbols():
    """Every function name declared in include/msda_b200.h (used by the symbol-export test)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", src)))


def lib():
    """Load (once) and return the CDLL; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ziragroundingdino_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or ziragroundingdino_b200/csrc/build.sh). There is no CPU or PyTorch fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.msda_b200_last_error.restype = ctypes.c_char_p
    L.msda_b200_launch_count.restype = _ll
    L.msda_b200_set_tuning.argtypes = [ctypes.c_char_p, _i]
    L.msda_b200_get_tuning.argtypes = [ctypes.c_char_p]
    for sfx in ("f32", "f64", "bf16", "f16"):
        getattr(L, "msda_forward_" + sfx).argtypes = _FWD
        getattr(L, "msda_backward_" + sfx).argtypes = _BWD
    L.msda_b200_gemm_last_error.restype = ctypes.c_char_p
    L.msda_linear_16.argtypes = [_vp, _vp, _vp, _ll, _i, _i, _vp, _i, _i, _vp, _i, _vp]
    L.msda_linear_accum_16.argtypes = [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp]
    L.msda_linear_accum2_16.argtypes = [_vp, _i, _vp, _i, _vp, _vp, _ll, _i, _vp, _vp, _i, _vp]
    L.msda_query_proj2_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp]
    L.msda_linear_add_layernorm_16.argtypes = [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _i, _vp]
    L.msda_linear_act_16.argtypes = [_vp, _vp, _vp, _ll, _i, _i, _vp, _i, _vp, _i, _vp]
    L.msda_linear_act_bits_16.argtypes = [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, _i, _vp]
    L.msda_query_proj_16.argtypes = [_vp, _vp, _vp, _vp, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp]
    L.msda_query_bwd_prep_16.argtypes = [_vp, _vp, _vp, _vp, _i, _vp, _ll, _i, _i, _i, _vp, _i, _i, _vp]
    L.msda_cast_mask_16.argtypes = [_vp, _vp, _ll, _i, _vp, _i, _vp]
    L.msda_zira_linear_16.argtypes = [_vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]
    L.msda_zira_bwd_prep_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _i, _vp]
    L.msda_add_layernorm_fwd_16.argtypes = [_vp, _vp, _vp, _vp, _ll, _i, ctypes.c_float, _vp, _vp, _vp, _vp, _i, _vp]
    L.msda_layernorm_fwd_16.argtypes = [_vp, _vp, _vp, _ll, _i, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _i, _vp]
    L.msda_add_layernorm_bwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _i, _vp]
    L.msda_group_norm_fwd_16.argtypes = [_vp, _ll, _vp, _vp, _i, _ll, _i, _i, ctypes.c_float, _vp, _ll, _vp, _vp, _i, _vp]
    L.msda_group_norm_bwd_16.argtypes = [_vp, _ll, _vp, _ll, _vp, _vp, _i, _ll, _i, _i, _vp, _ll, _vp, _i, _vp]
    L.msda_backward_fusedq_16.argtypes = [_vp] * 7 + [_i] * 8 + [_vp, _vp, _i, _i, _vp]
    L.msda_f16acc_scale.restype = ctypes.c_float
    L.msda_f16acc_scale.argtypes = [ctypes.c_uint32, _i]
    L.msda_grad_value_h16_rows.restype = _ll
    L.msda_grad_value_h16_rows.argtypes = [_vp, _i, _i]
    L.msda_backward_fusedq_h16.argtypes = [_vp] * 7 + [_i] * 8 + [_vp, _vp, _vp, _i, _vp]
    L.msda_cast_mask_h16.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _vp, _vp, _i, _i, _vp]
    L.msda_linear_f32.argtypes = [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp, _vp]
    L.msda_query_proj_f32.argtypes = [_vp, _vp, _vp, _vp, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp]
    L.msda_ffn_chain_fwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp]
    L.msda_ffn_chain_ln_fwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _i, _vp]
    L.msda_ffn_chain2_fwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _i, _vp]
    L.msda_ffn_chain2_ln_fwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _i, _vp]
    L.msda_ffn_chain2_bwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _i, _vp]
    L.msda_ffn_chain_bwd_16.argtypes = [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _i, _vp]
    L.msda_flatten_levels.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]
    L.msda_level_valid_counts.argtypes = [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]
    L.msda_encoder_proposals.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]
    L.msda_backward_workspace_bytes.restype = _ll
    L.msda_backward_workspace_bytes.argtypes = [_i, _i, _i]
    L.msda_backward_16_ws.argtypes = [_vp] * 7 + [_i] * 8 + [_vp] * 4 + [_i, _i, _vp, _ll, _vp]
    L.msda_query_post_f32.argtypes = [_vp, _vp, _i, _vp, _ll, _i, _i, _i, _vp, _vp, _vp]
    L.msda_biattn_splits.argtypes = [_i, _i]
    L.msda_biattn_pv_16.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]
    L.msda_biattn_combine_16.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]
    L.msda_biattn_ds_splits.argtypes = [_i, _i]
    L.msda_biattn_ds_16.argtypes = [_vp] * 6 + [_i] * 4 + [ctypes.c_float] + [_vp] * 8 + [_i, _i, _vp]
    L.msda_biattn_ds_terms_16.argtypes = [_vp] * 6 + [_i] * 4 + [ctypes.c_float] + [_vp] * 8 + [_i, _vp]
    L.msda_biattn_tn_splits.argtypes = [_i, _i]
    L.msda_biattn_tn_16.argtypes = [_vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp, _vp, _i, _i, _vp]
    L.msda_biattn_rowdot_16.argtypes = [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]
    L.msda_b200_probe_gather.argtypes = [_vp, _ll, _i, _i, _i, _vp, _vp]
    L.msda_b200_probe_scatter.argtypes = [_vp, _ll, _i, _i, _i, _vp]
    _lib = L
    # A/B knobs for measurement runs: MSDA_B200_TUNING="bwd_merge=1,fwd_passes=4"
    for item in filter(None, os.environ.get("MSDA_B200_TUNING", "").split(",")):
        k, _, v = item.partition("=")
        if L.msda_b200_set_tuning(k.strip().encode(), int(v)) != 0:
            raise RuntimeError("MSDA_B200_TUNING: unknown key %r" % k)
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().msda_b200_last_error().decode() or lib().msda_b200_gemm_last_error().decode()
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg))


def set_tuning(**kw):
    for k, v in kw.items():
        check(lib().msda_b200_set_tuning(k.encode(), int(v)), "msda_b200_set_tuning(%s)" % k)


def get_tuning(key):
    return lib().msda_b200_get_tuning(key.encode())


def launch_count():
    return int(lib().msda_b200_launch_count())
