"""Fused pieces of the encoder layer around MultiScaleDeformableAttention."""
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


_DERIVED = {}


def derived(p, kind):
    """Kernel-ready form of a parameter -- ``"f32"`` (fp32 contiguous copy: LayerNorm / GroupNorm affine, biases) or
    ``"t"`` (contiguous transpose: the dgrad operand) -- built once per parameter version instead of once per call
    (the layers around the path are frozen in the ZiRa configuration; each rebuild is a ~3 us kernel, ~60 per step).
    Nothing is cached while a CUDA graph is being captured (the copy would live in the graph's private pool).
    Like every version-keyed cache, a write through ``p.data`` is not seen: call ``clear_derived()`` after one."""
    key = (id(p), kind)
    ent = _DERIVED.get(key)
    inference = p.is_inference()          # no version counter: never cached
    if ent is not None and not inference and ent[0]() is p and ent[1] == (p._version, p.data_ptr(), p.dtype, p.device):
        return ent[2]
    with torch.no_grad():
        v = p.detach().float().contiguous() if kind == "f32" else p.detach().t().contiguous()
    if not inference and not (p.is_cuda and torch.cuda.is_current_stream_capturing()):
        if len(_DERIVED) > 1024:                      # entries of parameters that no longer exist
            for k in [k for k, e in _DERIVED.items() if e[0]() is None]:
                del _DERIVED[k]
        _DERIVED[key] = (weakref.ref(p), (p._version, p.data_ptr(), p.dtype, p.device), v)
    return v


def clear_derived():
    _DERIVED.clear()


class AddLayerNormFunction(Function):
    """``LayerNorm(x + r)`` over the last dimension, 16-bit activations, one kernel per direction
    (reference: ``src = self.norm1(src + self.dropout1(src2))``, transformer_for_adapter.py:901-902)."""

    @staticmethod
    def forward(ctx, x, r, weight, bias, eps):
        shape = x.shape
        C = shape[-1]
        x2, r2 = x.reshape(-1, C).contiguous(), r.reshape(-1, C).contiguous()
        R = x2.shape[0]
        z, y = torch.empty_like(x2), torch.empty_like(x2)
        mean = torch.empty(R, dtype=torch.float32, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        g32, b32 = derived(weight, "f32"), derived(bias, "f32")
        with torch.cuda.device(x.device):
            rc = _lib.lib().msda_add_layernorm_fwd_16(x2.data_ptr(), r2.data_ptr(), g32.data_ptr(), b32.data_ptr(), R, C, float(eps),
                                                      z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                      1 if x.dtype == torch.float16 else 0, _stream(x))
        _lib.check(rc, "msda_add_layernorm_fwd_16")
        ctx.save_for_backward(z, g32, mean, rstd)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, g32, mean, rstd = ctx.saved_tensors
        R, C = z.shape
        dy2 = dy.reshape(R, C).contiguous()
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            rc = _lib.lib().msda_add_layernorm_bwd_16(dy2.data_ptr(), z.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                      R, C, dz.data_ptr(), 1 if z.dtype == torch.float16 else 0, _stream(z))
        _lib.check(rc, "msda_add_layernorm_bwd_16")
        dw = db = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:   # affine parameters are frozen in the ZiRa configuration
            xhat = (z.float() - mean[:, None]) * rstd[:, None]
            dw = (dy2.float() * xhat).sum(0).to(dy.dtype) if ctx.needs_input_grad[2] else None
            db = dy2.float().sum(0).to(dy.dtype) if ctx.needs_input_grad[3] else None
        dz = dz.view(ctx.shape)
        return dz, dz, dw, db, None


def add_layer_norm(x, r, norm, dropout_p=0.0, training=False):
    """``norm(x + dropout(r))`` -- fused when possible (CUDA, 16-bit, no active dropout, C % 8 == 0, C <= 1024)."""
    C = x.shape[-1]
    if (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and r.dtype == x.dtype and not (training and dropout_p > 0)
            and C % 8 == 0 and C <= 1024 and norm.elementwise_affine and norm.bias is not None):
        return AddLayerNormFunction.apply(x, r, norm.weight, norm.bias, norm.eps)
    if training and dropout_p > 0:
        r = torch.nn.functional.dropout(r, dropout_p, True)
    return norm(x + r)


class GroupNormRowsFunction(Function):
    """``GroupNorm(G, C)`` over channels-last rows: x [N, HW, C] 16-bit -> y [N, HW, C] -- SURVEY.md 8(f) row N3, the
    reference's ``nn.GroupNorm(32, hidden_dim)`` of input_proj (groundingdino_dual_zero_rep_branch.py:258-277) on the
    layout the projection GEMM writes.  The incoming gradient may be a level slice of the flattened [N, S, C] gradient
    (image stride S*C): it is read in place."""

    @staticmethod
    def forward(ctx, x, weight, bias, G, eps):
        N, HW, C = x.shape
        x = x.contiguous()
        g32, b32 = derived(weight, "f32"), derived(bias, "f32")
        y = torch.empty_like(x)
        mean_rstd = torch.empty((N, G, 2), dtype=torch.float32, device=x.device)
        scratch = torch.zeros((N, C, 2), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().msda_group_norm_fwd_16(x.data_ptr(), HW * C, g32.data_ptr(), b32.data_ptr(), N, HW, C, G, float(eps),
                                                   y.data_ptr(), HW * C, mean_rstd.data_ptr(), scratch.data_ptr(),
                                                   1 if x.dtype == torch.float16 else 0, _stream(x))
        _lib.check(rc, "msda_group_norm_fwd_16")
        ctx.save_for_backward(x, g32, mean_rstd)
        ctx.G = G
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, g32, mean_rstd = ctx.saved_tensors
        N, HW, C = x.shape
        if not (dy.stride(2) == 1 and dy.stride(1) == C and dy.stride(0) % 8 == 0 and dy.data_ptr() % 16 == 0):
            dy = dy.contiguous()
        dx = torch.empty_like(x)
        scratch = torch.zeros((N, C, 2), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().msda_group_norm_bwd_16(dy.data_ptr(), dy.stride(0), x.data_ptr(), HW * C, g32.data_ptr(),
                                                   mean_rstd.data_ptr(), N, HW, C, ctx.G, dx.data_ptr(), HW * C,
                                                   scratch.data_ptr(), 1 if x.dtype == torch.float16 else 0, _stream(x))
        _lib.check(rc, "msda_group_norm_bwd_16")
        dw = scratch[:, :, 0].sum(0).to(dy.dtype) if ctx.needs_input_grad[1] else None
        db = scratch[:, :, 1].sum(0).to(dy.dtype) if ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


def group_norm_rows(x, norm):
    """GroupNorm over channels-last rows x [N, HW, C].  Fused on CUDA 16-bit; otherwise the library GroupNorm on the
    channels-first view (the reference's own call)."""
    C = x.shape[-1]
    if (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and norm.affine and C % 8 == 0 and C % norm.num_groups == 0
            and C // 8 <= 256 and 256 % (C // 8) == 0):
        return GroupNormRowsFunction.apply(x, norm.weight, norm.bias, norm.num_groups, norm.eps)
    return norm(x.transpose(1, 2)).transpose(1, 2)


def _linear_act16(x2d, w, bias_f32, relu=False, gate=None):
    R, K = x2d.shape
    Nout = w.shape[0]
    out = torch.empty((R, Nout), dtype=x2d.dtype, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_linear_act_16(x2d.data_ptr(), w.data_ptr(), 0 if bias_f32 is None else bias_f32.data_ptr(), R, K, Nout,
                                           out.data_ptr(), 1 if relu else 0, 0 if gate is None else gate.data_ptr(),
                                           1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_linear_act_16")
    return out


def _linear_act_bits16(x2d, w, bias_f32, relu_bits=None, gate_bits=None):
    """msda_linear_act_bits_16: bias + ReLU writing the 1-bit mask (relu_bits given, filled) or a gated store (gate_bits)."""
    R, K = x2d.shape
    Nout = w.shape[0]
    out = torch.empty((R, Nout), dtype=x2d.dtype, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_linear_act_bits_16(x2d.data_ptr(), w.data_ptr(), 0 if bias_f32 is None else bias_f32.data_ptr(), R, K,
                                                Nout, out.data_ptr(), 0 if relu_bits is None else relu_bits.data_ptr(),
                                                0 if gate_bits is None else gate_bits.data_ptr(),
                                                1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_linear_act_bits_16")
    return out


class FFN16Function(Function):
    """``linear2(relu(linear1(x)))`` on 16-bit activations (reference transformer_for_adapter.py:880-881).

    The two products with K = d_model run on the tcgen05 GEMM with the activation folded into the epilogue:
    forward ``h = relu(x W1^T + b1)`` (no separate ReLU pass) and backward ``dh = (dy W2) * (h > 0)`` (no separate
    threshold pass).  The two products with K = d_ffn (linear2 forward, linear1 dgrad) stay library GEMMs: their weight
    does not fit in shared memory and a single-CTA streaming kernel would lose to cuBLAS's 2-CTA multicast one."""

    # Measured on B200 (profiles/r1_gemm_ab.txt, R = 88 892): fused bias+ReLU forward 115 us vs cuBLAS + ReLU kernel 230 us;
    # gated dgrad 227 us vs cuBLAS + threshold kernel 250 us.  The switch exists for A/B runs.
    fuse_relu_backward = True
    # Frozen FFN weights (the ZiRa configuration): the backward needs relu'(h) only, so the forward keeps ONE BIT per hidden
    # activation ([d_ffn/32, R] words written by the same epilogue) and frees the 2048-wide activation: the gated dgrad then
    # streams its 364 MB output without reading 364 MB of h back (HBM moves ~3.9 TB/s per direction on this part).
    bit_gate_when_frozen = True

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        shape = x.shape
        x2d = x.reshape(-1, shape[-1]).contiguous()
        ctx.shape = shape
        ctx.bits = (FFN16Function.bit_gate_when_frozen and FFN16Function.fuse_relu_backward and not any(ctx.needs_input_grad[1:])
                    and w1.shape[0] % 32 == 0)
        if ctx.bits:
            from . import blocks      # the chained kernel's wrappers live beside the block Functions
            bits = torch.empty((w1.shape[0] // 32, x2d.shape[0]), dtype=torch.int32, device=x.device)
            ctx.chain = blocks.ffn_chain_ok(x2d.shape[1], w1.shape[0]) and w2.shape[0] == x2d.shape[1]
            if ctx.chain:     # linear1 -> ReLU -> linear2 in one launch, hidden activation kept on chip (csrc/layer_ffn_chain.cu)
                y = blocks.ffn_chain_fwd16(x2d, w1.contiguous(), derived(b1, "f32"), w2.contiguous(), derived(b2, "f32"), bits)
            else:
                h = _linear_act_bits16(x2d, w1.contiguous(), derived(b1, "f32"), relu_bits=bits)
                y = torch.nn.functional.linear(h, w2, b2)
            ctx.save_for_backward(bits, w1, w2)
            return y.view(*shape[:-1], w2.shape[0])
        h = _linear_act16(x2d, w1.contiguous(), derived(b1, "f32"), relu=True)
        y = torch.nn.functional.linear(h, w2, b2)
        ctx.save_for_backward(x2d, h, w1, w2)
        return y.view(*shape[:-1], w2.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        if ctx.bits:
            bits, w1, w2 = ctx.saved_tensors
            if not ctx.needs_input_grad[0]:
                return None, None, None, None, None
            if ctx.chain:
                from . import blocks
                dx = blocks.ffn_chain_bwd16(dy2, derived(w2, "t"), derived(w1, "t"), bits, accumulate=False)
                return dx.view(ctx.shape), None, None, None, None
            dh = _linear_act_bits16(dy2, derived(w2, "t"), None, gate_bits=bits)
            return (dh @ w1).view(ctx.shape), None, None, None, None
        x2d, h, w1, w2 = ctx.saved_tensors
        if FFN16Function.fuse_relu_backward:
            dh = _linear_act16(dy2, derived(w2, "t"), None, gate=h)    # (dy W2) gated by relu'(.) in the GEMM epilogue
        else:
            dh = torch.ops.aten.threshold_backward(dy2 @ w2, h, 0)
        dx = (dh @ w1).view(ctx.shape) if ctx.needs_input_grad[0] else None
        dw1 = dh.t() @ x2d if ctx.needs_input_grad[1] else None
        db1 = dh.float().sum(0).to(dy.dtype) if ctx.needs_input_grad[2] else None
        dw2 = dy2.t() @ h if ctx.needs_input_grad[3] else None
        db2 = dy2.float().sum(0).to(dy.dtype) if ctx.needs_input_grad[4] else None
        return dx, dw1, db1, dw2, db2


def ffn(x, linear1, linear2, dropout_p=0.0, training=False):
    """``linear2(dropout(relu(linear1(x))))`` -- fused when possible."""
    d_model, d_ffn = linear1.in_features, linear1.out_features
    if (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and not (training and dropout_p > 0) and d_model % 64 == 0
            and d_ffn % 32 == 0 and d_ffn <= 2048 and linear1.bias is not None and linear2.bias is not None):
        return FFN16Function.apply(x, linear1.weight, linear1.bias, linear2.weight, linear2.bias)
    h = torch.nn.functional.relu(linear1(x))
    if training and dropout_p > 0:
        h = torch.nn.functional.dropout(h, dropout_p, True)
    return linear2(h)
