"""Encoder-layer sub-blocks as single autograd Functions.

The reference's layer (transformer_for_adapter.py:890-907) uses ``src`` three times in the attention half -- residual,
``value``, and ``query = src + pos`` -- and twice in the FFN half, so plain autograd sums the incoming gradients with
separate elementwise add passes (3 per layer, 45 MB read twice + written once each).  Here each half is ONE Function
whose backward lets the dgrad GEMMs accumulate onto the gradient that is already there (``msda_linear_accum_16`` /
cuBLAS beta = 1), so no add pass remains.  Used only when everything inside the block is frozen (the ZiRa
configuration: gradients flow *through* the layers to the input-projection adapters); otherwise the layer composes the
stage-wise Functions as before.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _C, _lib, fused
from .layer_ops import _linear_act_bits16, _stream, derived


import os

# The two FFN products with K = d_ffn (linear2 forward, linear1 dgrad): library GEMMs by default -- their 1 MB weight
# does not stay resident in shared memory and the library's 2-CTA multicast kernel streams it at half the L2 traffic.
# MSDA_B200_OWN_K_FFN=1 (or blocks.OWN_K_FFN = True) runs them on the tcgen05 GEMM of csrc/proj_gemm.cu as well, so that
# no library kernel is left in the encoder layer (A/B in profiles/).
OWN_K_FFN = os.environ.get("MSDA_B200_OWN_K_FFN", "0") == "1"


def own_k_ffn():
    return OWN_K_FFN


# Default (MSDA_B200_FFN_CHAIN=0 / blocks.FFN_CHAIN = False turns it off): the whole FFN as ONE launch per direction with the hidden
# activation kept on chip (csrc/layer_ffn_chain.cu): linear1 -> ReLU (+ 1-bit mask) -> linear2 forward, and
# gate(dz W2) W1 + dz backward.  d_model = 256, d_ffn % 128 == 0.
FFN_CHAIN = os.environ.get("MSDA_B200_FFN_CHAIN", "1") == "1"


# The chained FFN on CTA pairs (tcgen05 cta_group::2, csrc/layer_ffn_chain2.cu): each SM streams half of every weight tile and
# the mid / final stages run on 16 warps.  Default on (bit-identical to the single-CTA kernel of csrc/layer_ffn_chain.cu,
# 15-20 % faster); MSDA_B200_FFN_PAIR=0 selects the single-CTA kernel.
FFN_PAIR = os.environ.get("MSDA_B200_FFN_PAIR", "1") == "1"


def _chain(name):
    return getattr(_lib.lib(), ("msda_ffn_chain2_" if FFN_PAIR else "msda_ffn_chain_") + name)


def ffn_chain_ok(C, F):
    return FFN_CHAIN and C == 256 and F % 128 == 0 and F <= 8192


def ffn_chain_fwd16(x2d, w1, b1_f32, w2, b2_f32, bits):
    R, C = x2d.shape
    F = w1.shape[0]
    out = torch.empty((R, C), dtype=x2d.dtype, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _chain("fwd_16")(x2d.data_ptr(), w1.data_ptr(), b1_f32.data_ptr(), w2.data_ptr(), b2_f32.data_ptr(), R, C, F,
                                              out.data_ptr(), bits.data_ptr(), 1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_ffn_chain_fwd_16")
    return out


# Default on: the residual + LayerNorm that follows the FFN runs in the chained kernel's final stage (MSDA_B200_FFN_LN=0 keeps the
# separate add+LayerNorm launch).
FFN_LN = os.environ.get("MSDA_B200_FFN_LN", "1") == "1"


def ffn_chain_ln_fwd16(x2d, w1, b1_f32, w2, b2_f32, g32, b32, eps, bits):
    """-> (z = x + ffn(x), y = LayerNorm(z), mean, rstd) in one launch."""
    R, C = x2d.shape
    F = w1.shape[0]
    z, y = torch.empty_like(x2d), torch.empty_like(x2d)
    mean = torch.empty(R, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty(R, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _chain("ln_fwd_16")(x2d.data_ptr(), w1.data_ptr(), b1_f32.data_ptr(), w2.data_ptr(), b2_f32.data_ptr(), R, C,
                                                 F, g32.data_ptr(), b32.data_ptr(), float(eps), z.data_ptr(), y.data_ptr(),
                                                 mean.data_ptr(), rstd.data_ptr(), bits.data_ptr(),
                                                 1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_ffn_chain_ln_fwd_16")
    return z, y, mean, rstd


def ffn_chain_bwd16(dz, w2_t, w1_t, bits, accumulate=True):
    """accumulate: dz <- dz + gate(dz W2) W1, in place (the residual path of the encoder block); otherwise returns
    gate(dz W2) W1 in a new tensor.  w2_t = W2^T [d_ffn, d_model], w1_t = W1^T [d_model, d_ffn]."""
    R, C = dz.shape
    F = w2_t.shape[0]
    out = dz if accumulate else torch.empty_like(dz)
    with torch.cuda.device(dz.device):
        rc = _chain("bwd_16")(dz.data_ptr(), w2_t.data_ptr(), w1_t.data_ptr(), bits.data_ptr(),
                                              dz.data_ptr() if accumulate else 0, R, C, F, out.data_ptr(),
                                              1 if dz.dtype == torch.float16 else 0, _stream(dz))
    _lib.check(rc, "msda_ffn_chain_bwd_16")
    return out


def _add_ln_fwd(x2, r2, g32, b32, eps):
    R, C = x2.shape
    z, y = torch.empty_like(x2), torch.empty_like(x2)
    mean = torch.empty(R, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(R, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        rc = _lib.lib().msda_add_layernorm_fwd_16(x2.data_ptr(), r2.data_ptr(), g32.data_ptr(), b32.data_ptr(), R, C, float(eps),
                                                  z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                  1 if x2.dtype == torch.float16 else 0, _stream(x2))
    _lib.check(rc, "msda_add_layernorm_fwd_16")
    return z, y, mean, rstd


def linear_add_ln16(x2d, w, b32_bias, residual, g32, b32, eps):
    """(z, y, mean, rstd) with z = residual + x2d @ w^T + bias and y = LayerNorm(z): the projection, the residual add and the
    LayerNorm as ONE GEMM launch (msda_linear_add_layernorm_16; the statistics come from the rounded z it stores)."""
    R, K = x2d.shape
    Nout = w.shape[0]
    assert x2d.is_contiguous() and w.is_contiguous() and residual.is_contiguous() and residual.shape == (R, Nout)
    z, y = torch.empty_like(residual), torch.empty_like(residual)
    mean = torch.empty(R, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty(R, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_linear_add_layernorm_16(x2d.data_ptr(), w.data_ptr(), 0 if b32_bias is None else b32_bias.data_ptr(), R, K,
                                                     Nout, residual.data_ptr(), g32.data_ptr(), b32.data_ptr(), float(eps),
                                                     z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                     1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_linear_add_layernorm_16")
    return z, y, mean, rstd


def _add_ln_bwd(dy2, z, g32, mean, rstd):
    R, C = z.shape
    dz = torch.empty_like(z)
    with torch.cuda.device(z.device):
        rc = _lib.lib().msda_add_layernorm_bwd_16(dy2.data_ptr(), z.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                  R, C, dz.data_ptr(), 1 if z.dtype == torch.float16 else 0, _stream(z))
    _lib.check(rc, "msda_add_layernorm_bwd_16")
    return dz


def linear_accum16(x2d, w, accum, out=None):
    """out = accum + x2d @ w^T (16-bit); ``out`` defaults to accumulating in place."""
    R, K = x2d.shape
    Nout = w.shape[0]
    out = accum if out is None else out
    assert x2d.is_contiguous() and w.is_contiguous() and accum.is_contiguous() and accum.shape == (R, Nout) and accum.dtype == x2d.dtype
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_linear_accum_16(x2d.data_ptr(), w.data_ptr(), 0, R, K, Nout, accum.data_ptr(), out.data_ptr(),
                                             1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_linear_accum_16")
    return out


def linear_accum2_16(x1, x2, w12, accum, out=None):
    """out = accum + [x1 | x2] @ w12^T (16-bit) without concatenating the activations: ONE GEMM whose k-blocks come from two
    operands (msda_linear_accum2_16); ``out`` defaults to accumulating in place."""
    R, K1 = x1.shape
    K2 = x2.shape[1]
    Nout = w12.shape[0]
    out = accum if out is None else out
    assert x1.is_contiguous() and x2.is_contiguous() and w12.is_contiguous() and accum.is_contiguous()
    assert x2.shape[0] == R and w12.shape[1] == K1 + K2 and accum.shape == (R, Nout) and accum.dtype == x1.dtype == x2.dtype
    with torch.cuda.device(x1.device):
        rc = _lib.lib().msda_linear_accum2_16(x1.data_ptr(), K1, x2.data_ptr(), K2, w12.data_ptr(), 0, R, Nout, accum.data_ptr(),
                                              out.data_ptr(), 1 if x1.dtype == torch.float16 else 0, _stream(x1))
    _lib.check(rc, "msda_linear_accum2_16")
    return out


# A/B switches (tests, measurements).
# FOLD_POS (default on): `query = src + pos` as a second operand of the query projection instead of an elementwise pass --
#   in the captured step the projection grows by 16 us and the 25 us add disappears (config 2: -54 us, config 4: -85 us),
#   and the locations come from the exact sum (7e-7 from fp64 instead of 2e-3 for the rounded 16-bit sum).
# DGRAD_CAT (default OFF, a measured loss): the block's two input dgrads as one K-concatenated product (K = 640).  It saves
#   one read + write of the accumulator, but W no longer fits beside the ring at 256 columns, so two CTAs stream every
#   activation tile: 62 us against 2 x 24 us at 4 images (config 2 +83 us, config 4 +213 us per step).
# OUT_LN (default on): `norm1(src + output_proj(core))` as one GEMM launch with the LayerNorm in its epilogue
#   (msda_linear_add_layernorm_16) instead of the projection + the add+LayerNorm kernel: 49.9 us against 25.0 + 32.9 us at 4
#   images (this epilogue makes two passes over the accumulator and issues four dependent stores per warp and tile, so it,
#   not HBM, bounds the launch: the saving is most of the 91 MB round trip of the projection's output, not a whole kernel).
OUT_LN = os.environ.get("MSDA_B200_OUT_LN", "1") != "0"
FOLD_POS = os.environ.get("MSDA_B200_FOLD_POS", "1") != "0"
DGRAD_CAT = os.environ.get("MSDA_B200_DGRAD_CAT", "0") == "1"


class SelfAttnBlockFunction(Function):
    """``norm1(src + MSDeformAttn(query = src + pos, value = src))`` with every parameter frozen: returns the block
    output; backward returns d(src) only, with the value-projection and query-projection dgrads accumulated onto the
    residual gradient inside the GEMM epilogues."""

    @staticmethod
    def forward(ctx, src, pos, row_mask, reference_points, spatial_shapes, level_start_index, prep, M, L, P, im2col_step,
                g32, b32, eps):
        N, S, C = src.shape
        src2d = src.reshape(N * S, C).contiguous()
        ref = reference_points.to(torch.float32).contiguous()
        ref_dim = ref.shape[-1]
        value = fused.linear16(src2d, prep.w_v, prep.b_v, row_mask).view(N, S, M, C // M)
        if pos is not None and FOLD_POS and pos.shape == src.shape and pos.dtype == src.dtype:
            # (src + pos) W^T = src W^T + pos W^T, both products accumulated in fp32 by the one GEMM (transformer_for_adapter.py:867-869, :893)
            loc, aw = fused.query_proj16(src2d, prep.w_cat, prep.b_cat, ref, ref_dim, spatial_shapes, M, L, P,
                                         q_add=pos.reshape(N * S, C).contiguous())
        else:
            q2d = src2d if pos is None else (src + pos).reshape(N * S, C)
            loc, aw = fused.query_proj16(q2d, prep.w_cat, prep.b_cat, ref, ref_dim, spatial_shapes, M, L, P)
        loc, aw = loc.view(N, S, M, L, P, 2), aw.view(N, S, M, L, P)
        core = _C.ms_deform_attn_forward(value, spatial_shapes, level_start_index, loc, aw, im2col_step)
        if OUT_LN and C in (128, 256):
            z, y, mean, rstd = linear_add_ln16(core.view(N * S, C), prep.w_o, prep.b_o, src2d, g32, b32, eps)
        else:
            attn = fused.linear16(core.view(N * S, C), prep.w_o, prep.b_o)
            z, y, mean, rstd = _add_ln_fwd(src2d, attn, g32, b32, eps)
        ctx.dims = (N, S, C, M, L, P, ref_dim, im2col_step)
        ctx.prep = prep
        ctx.save_for_backward(row_mask, ref, spatial_shapes, level_start_index, value, loc, aw, z, g32, mean, rstd)
        return y.view(N, S, C)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        row_mask, ref, spatial_shapes, level_start_index, value, loc, aw, z, g32, mean, rstd = ctx.saved_tensors
        N, S, C, M, L, P, ref_dim, im2col_step = ctx.dims
        prep = ctx.prep
        dt = value.dtype
        dz = _add_ln_bwd(dy.reshape(N * S, C).contiguous(), z, g32, mean, rstd)     # d(src) via the residual AND d(attn out)
        d_core = fused.linear16(dz, prep.w_o_t)
        if fused.fusedq_ok(M, L, P, C // M):
            gv16, dq_cat = fused.backward_fusedq_gv16(value, spatial_shapes, level_start_index, loc, aw, d_core, ref, ref_dim,
                                                      row_mask)
        else:
            grad_value, grad_loc, grad_aw = _C.ms_deform_attn_backward(value, spatial_shapes, level_start_index, loc, aw,
                                                                       d_core.view(N, S, C), im2col_step)
            dq_cat = fused.query_bwd_prep16(grad_loc, grad_aw, aw, ref, ref_dim, spatial_shapes, N * S, M, L, P, dt)
            gv16 = fused.cast_mask16(grad_value.view(N * S, C), row_mask, dt)
        if DGRAD_CAT and dq_cat.shape[1] % 64 == 0 and dq_cat.shape[1] == prep.w_cat_t.shape[1]:
            # dz += d(value_proj input) + d(query) (= d(src) through query = src + pos): one product over [grad_value | dq]
            d_src = linear_accum2_16(gv16, dq_cat, prep.w_vq_t, dz)
        else:
            d_src = linear_accum16(gv16, prep.w_v_t, dz)            # dz += d(value_proj input)
            d_src = linear_accum16(dq_cat, prep.w_cat_t, d_src)     # dz += d(query)
        return (d_src.view(N, S, C),) + (None,) * 13


class FFNBlockFunction(Function):
    """``norm2(x + linear2(relu(linear1(x))))`` with frozen weights: 1-bit ReLU mask, and the linear1 dgrad accumulates
    onto the residual gradient (cuBLAS beta = 1) instead of a separate add."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, g32, b32, eps):
        shape = x.shape
        x2d = x.reshape(-1, shape[-1]).contiguous()
        bits = torch.empty((w1.shape[0] // 32, x2d.shape[0]), dtype=torch.int32, device=x.device)
        ctx.chain = ffn_chain_ok(x2d.shape[1], w1.shape[0])
        if ctx.chain:
            if FFN_LN:
                z, y, mean, rstd = ffn_chain_ln_fwd16(x2d, w1.contiguous(), derived(b1, "f32"), w2.contiguous(), derived(b2, "f32"),
                                                      g32, b32, eps, bits)
            else:
                y2 = ffn_chain_fwd16(x2d, w1.contiguous(), derived(b1, "f32"), w2.contiguous(), derived(b2, "f32"), bits)
                z, y, mean, rstd = _add_ln_fwd(x2d, y2, g32, b32, eps)
            ctx.save_for_backward(bits, w1, w2, z, g32, mean, rstd)
            ctx.shape = shape
            return y.view(shape)
        h = _linear_act_bits16(x2d, w1.contiguous(), derived(b1, "f32"), relu_bits=bits)
        if own_k_ffn():      # K = d_ffn on the tcgen05 GEMM too (W2 streams through the ring): no library kernel in the block
            y2 = fused.linear16(h, w2.contiguous(), derived(b2, "f32"))
        else:
            y2 = torch.nn.functional.linear(h, w2, b2)
        z, y, mean, rstd = _add_ln_fwd(x2d, y2, g32, b32, eps)
        ctx.save_for_backward(bits, w1, w2, z, g32, mean, rstd)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        bits, w1, w2, z, g32, mean, rstd = ctx.saved_tensors
        dz = _add_ln_bwd(dy.reshape(z.shape).contiguous(), z, g32, mean, rstd)
        if ctx.chain:
            return (ffn_chain_bwd16(dz, derived(w2, "t"), derived(w1, "t"), bits).view(ctx.shape),) + (None,) * 7
        dh = _linear_act_bits16(dz, derived(w2, "t"), None, gate_bits=bits)
        if own_k_ffn():
            dx = linear_accum16(dh, derived(w1, "t"), dz)      # dz += dh W1, accumulated in the GEMM epilogue
        else:
            dx = dz.addmm_(dh, w1)      # in place (beta = 1): the out-of-place form first copies dz into its output (a 45 MB memcpy)
        return (dx.view(ctx.shape),) + (None,) * 7


def _frozen(*mods):
    return not any(p.requires_grad for m in mods for p in m.parameters())


def _ln_ok(norm, C):
    return norm.elementwise_affine and norm.bias is not None and C % 8 == 0 and C <= 1024


def self_attn_block(layer, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask):
    """The attention half of an encoder layer as one Function, or None when the block form does not apply."""
    attn, norm = layer.self_attn, layer.norm1
    C = src.shape[-1]
    if not (src.is_cuda and src.dtype in (torch.bfloat16, torch.float16) and attn.batch_first and _ln_ok(norm, C)
            and _frozen(attn, norm) and torch.is_grad_enabled() and src.requires_grad
            and (pos is None or (not pos.requires_grad and pos.dtype == src.dtype))
            and not (layer.training and getattr(layer.dropout1, "p", 0.0) > 0)
            and attn.value_proj_adapter is None and attn.output_proj_adapter is None
            and attn._use_fused(src, reference_points)):
        return None
    prep = attn._prepared()
    row_mask = None if key_padding_mask is None else key_padding_mask.reshape(-1).to(torch.uint8).contiguous()
    attn.zero_inter_loss = None
    return SelfAttnBlockFunction.apply(src, pos, row_mask, reference_points, spatial_shapes, level_start_index, prep,
                                       attn.num_heads, attn.num_levels, attn.num_points, attn.im2col_step,
                                       derived(norm.weight, "f32"), derived(norm.bias, "f32"), norm.eps)


def ffn_block(layer, x):
    """The FFN half of an encoder layer as one Function, or None when the block form does not apply."""
    l1, l2, norm = layer.linear1, layer.linear2, layer.norm2
    C = x.shape[-1]
    if not (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and _ln_ok(norm, C) and _frozen(l1, l2, norm)
            and torch.is_grad_enabled() and x.requires_grad and l1.bias is not None and l2.bias is not None
            and C % 64 == 0 and l1.out_features % 32 == 0 and l1.out_features <= 2048
            and not (layer.training and (getattr(layer.dropout2, "p", 0.0) > 0 or getattr(layer.dropout3, "p", 0.0) > 0))):
        return None
    return FFNBlockFunction.apply(x, l1.weight, l1.bias, l2.weight, l2.bias, derived(norm.weight, "f32"),
                                  derived(norm.bias, "f32"), norm.eps)
