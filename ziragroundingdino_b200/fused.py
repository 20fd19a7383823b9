"""Fused 16-bit execution of ``MultiScaleDeformableAttention.forward`` (reference ms_deform_attn.py:281-350).

Five launches replace the reference's ~20 eager ops per call:

  value_proj + key-padding mask ............ msda_linear_16        (tcgen05 GEMM, mask in the epilogue)
  sampling_offsets | attention_weights ..... msda_query_proj_16    (ONE tcgen05 GEMM; epilogue = sampling
                                             locations + softmax, fp32 out)
  bilinear gather .......................... msda_forward_{bf16,f16}
  output_proj .............................. msda_linear_16

and the backward is: dgrad GEMM, scatter kernel, one elementwise prep kernel (softmax / location
backward + down-cast), dgrad GEMM (K = 3*M*L*P), mask+cast kernel, dgrad GEMM.  Weight gradients, only
needed when the module's own linears are trainable (they are frozen in the ZiRa configuration,
groundingdino/config/GroundingDINO_SwinT_OGC_rep.py:50), are plain library GEMMs.
"""
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _C, _lib


GEMM_MAX_N = 2048   # pg::MAX_N of csrc/proj_gemm.cu: widest output one launch of the tcgen05 GEMM covers


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def linear16(x2d, w, bias_f32=None, row_mask=None, out_f32=False):
    """x2d [R, K] (bf16/f16, contiguous) @ w[Nout, K]^T + bias -> [R, Nout]."""
    R, K = x2d.shape
    Nout = w.shape[0]
    assert x2d.is_contiguous() and w.is_contiguous() and w.dtype == x2d.dtype and w.shape[1] == K
    out = torch.empty((R, Nout), dtype=torch.float32 if out_f32 else x2d.dtype, device=x2d.device)
    if R == 0:
        return out
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_linear_16(x2d.data_ptr(), w.data_ptr(), 0 if bias_f32 is None else bias_f32.data_ptr(), R, K,
                                       Nout, out.data_ptr(), Nout, 1 if out_f32 else 0,
                                       0 if row_mask is None else row_mask.data_ptr(),
                                       1 if x2d.dtype == torch.float16 else 0, _stream(x2d))
    _lib.check(rc, "msda_linear_16")
    return out


def query_proj16(q2d, w_cat, bias_cat_f32, ref, ref_dim, spatial_shapes, M, L, P, q_add=None):
    """-> (loc [R, M, L, P, 2] fp32, aw [R, M, L, P] fp32).  ``q_add`` [R, K]: the projection of ``q2d + q_add`` without
    forming the sum (two activation operands against one resident copy of the weight: x1 W^T + x2 W^T in fp32)."""
    R, K = q2d.shape
    loc = torch.empty((R, M, L, P, 2), dtype=torch.float32, device=q2d.device)
    aw = torch.empty((R, M, L, P), dtype=torch.float32, device=q2d.device)
    if R == 0:
        return loc, aw
    assert q_add is None or (q_add.shape == q2d.shape and q_add.dtype == q2d.dtype and q_add.is_contiguous())
    with torch.cuda.device(q2d.device):
        rc = _lib.lib().msda_query_proj2_16(q2d.data_ptr(), 0 if q_add is None else q_add.data_ptr(), w_cat.data_ptr(),
                                            bias_cat_f32.data_ptr(), ref.data_ptr(), ref_dim, spatial_shapes.data_ptr(), R, K, M,
                                            L, P, loc.data_ptr(), aw.data_ptr(), 1 if q2d.dtype == torch.float16 else 0,
                                            _stream(q2d))
    _lib.check(rc, "msda_query_proj2_16")
    return loc, aw


def padded_k(n):
    """Row length of the stacked query gradient: 3*M*L*P rounded up to the GEMM's K granularity (64)."""
    return (n + 63) // 64 * 64


def query_bwd_prep16(grad_loc, grad_aw, aw, ref, ref_dim, spatial_shapes, R, M, L, P, dtype):
    ld = padded_k(3 * M * L * P)
    out = torch.empty((R, ld), dtype=dtype, device=aw.device)
    with torch.cuda.device(aw.device):
        rc = _lib.lib().msda_query_bwd_prep_16(grad_loc.data_ptr(), grad_aw.data_ptr(), aw.data_ptr(), ref.data_ptr(), ref_dim,
                                               spatial_shapes.data_ptr(), R, M, L, P, out.data_ptr(), ld,
                                               1 if dtype == torch.float16 else 0, _stream(aw))
    _lib.check(rc, "msda_query_bwd_prep_16")
    return out


def cast_mask16(x_f32_2d, row_mask, dtype):
    rows, cols = x_f32_2d.shape
    out = torch.empty((rows, cols), dtype=dtype, device=x_f32_2d.device)
    with torch.cuda.device(x_f32_2d.device):
        rc = _lib.lib().msda_cast_mask_16(x_f32_2d.data_ptr(), 0 if row_mask is None else row_mask.data_ptr(), rows, cols,
                                          out.data_ptr(), 1 if dtype == torch.float16 else 0, _stream(x_f32_2d))
    _lib.check(rc, "msda_cast_mask_16")
    return out


def backward_fusedq16(value, spatial_shapes, level_start_index, loc, aw, grad_core, ref, ref_dim):
    """Scatter backward with the query-side epilogue backward fused in -> (grad_value fp32, dq_cat 16-bit)."""
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    grad_value = torch.empty(value.shape, dtype=torch.float32, device=value.device)
    dq = torch.empty((N * Lq, 3 * M * L * P), dtype=value.dtype, device=value.device)
    ws = _C.backward_workspace(value, N, M, D, Lq, P)
    with torch.cuda.device(value.device):
        if ws is not None:
            rc = _lib.lib().msda_backward_16_ws(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                loc.data_ptr(), aw.data_ptr(), grad_core.data_ptr(), ref.data_ptr(), ref_dim, N, S,
                                                M, D, L, Lq, P, grad_value.data_ptr(), 0, 0, dq.data_ptr(), 1,
                                                1 if value.dtype == torch.float16 else 0, ws.data_ptr(), ws.numel() * 8,
                                                _stream(value))
        else:
            rc = _lib.lib().msda_backward_fusedq_16(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                    loc.data_ptr(), aw.data_ptr(), grad_core.data_ptr(), ref.data_ptr(), ref_dim, N,
                                                    S, M, D, L, Lq, P, grad_value.data_ptr(), dq.data_ptr(), 1,
                                                    1 if value.dtype == torch.float16 else 0, _stream(value))
    _lib.check(rc, "msda_backward_fusedq_16")
    return grad_value, dq


_H16_ROWS = {}


def h16_rows(spatial_shapes, Lq, with_host_array=False):
    """Rows per image of the scaled-fp16 grad_value map (levels with their replicas) for these shapes."""
    import ctypes
    from .ms_deform_attn import _host_shapes
    hs = _host_shapes(spatial_shapes)
    hit = _H16_ROWS.get((hs, Lq))
    if hit is None:
        arr = (ctypes.c_int64 * (2 * len(hs)))(*[d for hw in hs for d in hw])
        rows = int(_lib.lib().msda_grad_value_h16_rows(arr, len(hs), Lq))
        if rows <= 0:
            raise RuntimeError("msda_grad_value_h16_rows: bad shapes %r" % (hs,))
        hit = _H16_ROWS[(hs, Lq)] = (rows, arr)
    return hit if with_host_array else hit[0]


def backward_fusedq_h16(value, spatial_shapes, level_start_index, loc, aw, grad_core, ref, ref_dim):
    """The same backward with grad_value accumulated in SCALED fp16 (include/msda_b200.h: msda_backward_fusedq_h16): returns
    (the map buffer -- N * rows_h * M * D halves and a 128-byte tail holding max |grad_core| -- as an int16 tensor, dq_cat)."""
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    rows_h, shapes_host = h16_rows(spatial_shapes, Lq, with_host_array=True)
    buf = torch.empty((N * rows_h * M * D + 64,), dtype=torch.int16, device=value.device)
    dq = torch.empty((N * Lq, 3 * M * L * P), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.lib().msda_backward_fusedq_h16(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                 loc.data_ptr(), aw.data_ptr(), grad_core.data_ptr(), ref.data_ptr(), ref_dim, N, S,
                                                 M, D, L, Lq, P, buf.data_ptr(), shapes_host, dq.data_ptr(),
                                                 1 if value.dtype == torch.float16 else 0, _stream(value))
    _lib.check(rc, "msda_backward_fusedq_h16")
    return buf, dq


def cast_mask_h16(buf, spatial_shapes, level_start_index, row_mask, N, S, cols, Lq, dtype):
    """Scaled-fp16 grad_value map -> [N*S, cols] in `dtype` (16-bit storage or fp32): replicas summed, un-scaled, padded
    rows zeroed."""
    out = torch.empty((N * S, cols), dtype=dtype, device=buf.device)
    rows_h = h16_rows(spatial_shapes, Lq)
    assert buf.numel() == N * rows_h * cols + 64
    with torch.cuda.device(buf.device):
        rc = _lib.lib().msda_cast_mask_h16(buf.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                           spatial_shapes.shape[0], N, S, cols, Lq, rows_h,
                                           0 if row_mask is None else row_mask.data_ptr(), out.data_ptr(),
                                           1 if dtype == torch.float32 else 0, 1 if dtype == torch.float16 else 0, _stream(buf))
    _lib.check(rc, "msda_cast_mask_h16")
    return out


# Opt-in (MSDA_B200_F16ACC=1 or this switch): grad_value of the fused 16-bit modules accumulates in scaled fp16 -- half the
# reduction bytes, ~1.3e-3 rms / 3.5e-3 max of max |grad_value| against fp32 accumulation; measured -7 % on the scatter at
# config 2.  The default keeps fp32 accumulation.
f16_accumulate = os.environ.get("MSDA_B200_F16ACC", "0") == "1"


def backward_fusedq_gv16(value, spatial_shapes, level_start_index, loc, aw, grad_core, ref, ref_dim, row_mask):
    """Fused-query backward -> (grad_value as masked 16-bit rows [N*S, M*D], dq_cat): the operands of the two dgrad GEMMs."""
    N, S, M, D = value.shape
    if f16_accumulate:
        buf, dq = backward_fusedq_h16(value, spatial_shapes, level_start_index, loc, aw, grad_core, ref, ref_dim)
        return cast_mask_h16(buf, spatial_shapes, level_start_index, row_mask, N, S, M * D, loc.shape[1], value.dtype), dq
    grad_value, dq = backward_fusedq16(value, spatial_shapes, level_start_index, loc, aw, grad_core, ref, ref_dim)
    return cast_mask16(grad_value.view(N * S, M * D), row_mask, value.dtype), dq


fuse_query_backward = True   # A/B switch (tests, benchmarks)


def fusedq_ok(M, L, P, D):
    """The scatter kernel with the query-side backward fused in writes dq rows of exactly 3*M*L*P columns (L = P = 4,
    D = 32); the dgrad GEMM that reads them needs K % 64 == 0, so M % 4 != 0 (e.g. 192 / 6 heads) takes the un-fused
    pair msda_backward + msda_query_bwd_prep_16, whose output row IS padded."""
    return fuse_query_backward and (L, P, D) == (4, 4, 32) and (3 * M * L * P) % 64 == 0


def supported(embed_dim, M, L, P, dtype):
    lp = L * P
    tpg = lp // 4
    pow2 = tpg > 0 and (tpg & (tpg - 1)) == 0
    return (dtype in (torch.bfloat16, torch.float16) and embed_dim % 64 == 0 and embed_dim <= 1024 and lp % 4 == 0
            and (not pow2 or 32 % tpg == 0) and (M * lp) % 32 == 0 and padded_k(3 * M * lp) <= 2048 and L <= 16)


class Prepared:
    """Weights of one module in the form the kernels consume (built once per parameter version):
    contiguous 16-bit matrices and their transposes for the dgrad products, the stacked
    [sampling_offsets; attention_weights] matrix, fp32 biases."""

    def __init__(self, w_v, b_v, w_off, b_off, w_aw, b_aw, w_o, b_o):
        with torch.no_grad():
            self.w_v, self.w_o = w_v.detach().contiguous(), w_o.detach().contiguous()
            self.w_cat = torch.cat([w_off.detach(), w_aw.detach()], 0).contiguous()
            self.b_v, self.b_o = b_v.detach().float(), b_o.detach().float()
            self.b_cat = torch.cat([b_off.detach(), b_aw.detach()], 0).float()
            self.w_v_t, self.w_o_t = self.w_v.t().contiguous(), self.w_o.t().contiguous()
            n = self.w_cat.shape[0]
            self.w_cat_t = torch.zeros((self.w_cat.shape[1], padded_k(n)), dtype=self.w_cat.dtype, device=self.w_cat.device)
            self.w_cat_t[:, :n] = self.w_cat.t()          # zero-padded along K to the GEMM's granularity
            # [W_v^T | W_cat^T]: the value-projection and query-projection dgrads as ONE product over [grad_value | dq]
            self.w_vq_t = torch.cat([self.w_v_t, self.w_cat_t], 1).contiguous()


class FusedMSDeformAttnFunction(Function):
    """Whole-module forward/backward on 16-bit activations. Inputs are batch-first and contiguous.
    ``prep`` carries the kernel-ready weights; the eight raw parameters are passed only so autograd can
    route weight gradients to them when they are trainable."""

    @staticmethod
    def forward(ctx, query, value_in, row_mask, reference_points, spatial_shapes, level_start_index, prep, M, L, P,
                im2col_step, w_v, b_v, w_off, b_off, w_aw, b_aw, w_o, b_o):
        N, Lq, C = query.shape
        S = value_in.shape[1]
        q2d = query.reshape(N * Lq, C)
        v2d = value_in.reshape(N * S, C)
        ref = reference_points.to(torch.float32).contiguous()
        ref_dim = ref.shape[-1]
        value = linear16(v2d, prep.w_v, prep.b_v, row_mask).view(N, S, M, C // M)
        loc, aw = query_proj16(q2d, prep.w_cat, prep.b_cat, ref, ref_dim, spatial_shapes, M, L, P)
        loc = loc.view(N, Lq, M, L, P, 2)
        aw = aw.view(N, Lq, M, L, P)
        core = _C.ms_deform_attn_forward(value, spatial_shapes, level_start_index, loc, aw, im2col_step)
        out = linear16(core.view(N * Lq, C), prep.w_o, prep.b_o).view(N, Lq, C)
        ctx.dims = (N, Lq, S, C, M, L, P, ref_dim, im2col_step)
        ctx.prep = prep
        ctx.wgrad = any(ctx.needs_input_grad[11:19])
        ctx.save_for_backward(query if ctx.wgrad else None, value_in if ctx.wgrad else None, row_mask, ref, spatial_shapes,
                              level_start_index, value, loc, aw, core if ctx.wgrad else None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        query, value_in, row_mask, ref, spatial_shapes, level_start_index, value, loc, aw, core = ctx.saved_tensors
        N, Lq, S, C, M, L, P, ref_dim, im2col_step = ctx.dims
        prep = ctx.prep
        dt = value.dtype
        g2d = grad_out.contiguous().view(N * Lq, C)
        d_core = linear16(g2d, prep.w_o_t)
        if fusedq_ok(M, L, P, C // M):
            gv16, dq_cat = backward_fusedq_gv16(value, spatial_shapes, level_start_index, loc, aw, d_core, ref, ref_dim, row_mask)
        else:
            grad_value, grad_loc, grad_aw = _C.ms_deform_attn_backward(value, spatial_shapes, level_start_index, loc, aw,
                                                                       d_core.view(N, Lq, C), im2col_step)
            dq_cat = query_bwd_prep16(grad_loc, grad_aw, aw, ref, ref_dim, spatial_shapes, N * Lq, M, L, P, dt)
            gv16 = cast_mask16(grad_value.view(N * S, C), row_mask, dt)
        d_query = linear16(dq_cat, prep.w_cat_t).view(N, Lq, C) if ctx.needs_input_grad[0] else None
        d_value_in = linear16(gv16, prep.w_v_t).view(N, S, C) if ctx.needs_input_grad[1] else None
        grads = [None] * 8
        if ctx.wgrad:  # plain library GEMMs; the module's own linears are frozen in the ZiRa configuration
            n_loc = 2 * M * L * P
            q2d, v2d = query.reshape(N * Lq, C), value_in.reshape(N * S, C)
            n_cat = 3 * M * L * P
            dw_cat = dq_cat[:, :n_cat].t() @ q2d
            db_cat = dq_cat[:, :n_cat].float().sum(0)
            grads = [gv16.t() @ v2d, gv16.float().sum(0).to(dt), dw_cat[:n_loc], db_cat[:n_loc].to(dt), dw_cat[n_loc:],
                     db_cat[n_loc:].to(dt), g2d.t() @ core.view(N * Lq, C), g2d.float().sum(0).to(dt)]
            grads = [g if ctx.needs_input_grad[11 + i] else None for i, g in enumerate(grads)]
        return (d_query, d_value_in, None, None, None, None, None, None, None, None, None, *grads)


# ---------------------------------------------------------------------------------------------------
# Stage-wise autograd Functions: used when the module cannot run as ONE fused Function, i.e. with
# un-merged ZiRa branches in training mode (their GEMM has its own epilogue, loss output and backward).
# ---------------------------------------------------------------------------------------------------
class Linear16Function(Function):
    """y = x W^T + b with an optional row mask (rows zeroed), x [R, K] 16-bit."""

    @staticmethod
    def forward(ctx, x2d, w, b, row_mask):
        ctx.save_for_backward(x2d, w, row_mask)
        return linear16(x2d, w.contiguous(), None if b is None else b.float(), row_mask)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2d, w, row_mask = ctx.saved_tensors
        gy = gy.contiguous()
        if row_mask is not None:
            gy = gy.masked_fill(row_mask.bool()[:, None], 0)
        gx = linear16(gy, w.t().contiguous()) if ctx.needs_input_grad[0] else None
        gw = gy.t() @ x2d if ctx.needs_input_grad[1] else None
        gb = gy.float().sum(0).to(gy.dtype) if ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


class QueryProj16Function(Function):
    """(query, W_off, b_off, W_aw, b_aw, ref) -> (sampling locations, softmax weights), both fp32."""

    @staticmethod
    def forward(ctx, q2d, w_off, b_off, w_aw, b_aw, ref, spatial_shapes, M, L, P):
        w_cat = torch.cat([w_off, w_aw], 0).contiguous()
        b_cat = torch.cat([b_off, b_aw], 0).float()
        ref = ref.to(torch.float32).contiguous()
        loc, aw = query_proj16(q2d, w_cat, b_cat, ref, ref.shape[-1], spatial_shapes, M, L, P)
        ctx.save_for_backward(q2d, w_cat, ref, spatial_shapes, aw)
        ctx.dims = (M, L, P)
        return loc, aw

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loc, g_aw):
        q2d, w_cat, ref, spatial_shapes, aw = ctx.saved_tensors
        M, L, P = ctx.dims
        R = q2d.shape[0]
        dq_cat = query_bwd_prep16(g_loc.contiguous(), g_aw.contiguous(), aw, ref, ref.shape[-1], spatial_shapes, R, M, L, P,
                                  q2d.dtype)
        n_cat = 3 * M * L * P
        gq = None
        if ctx.needs_input_grad[0]:
            w_t = torch.zeros((w_cat.shape[1], padded_k(n_cat)), dtype=w_cat.dtype, device=w_cat.device)
            w_t[:, :n_cat] = w_cat.t()
            gq = linear16(dq_cat, w_t)
        n_loc = 2 * M * L * P
        gw = gb = None
        if any(ctx.needs_input_grad[1:5]):
            gw = dq_cat[:, :n_cat].t() @ q2d
            gb = dq_cat[:, :n_cat].float().sum(0).to(q2d.dtype)
        pick = lambda i, t: t if (t is not None and ctx.needs_input_grad[i]) else None
        return (gq, pick(1, None if gw is None else gw[:n_loc]), pick(2, None if gb is None else gb[:n_loc]),
                pick(3, None if gw is None else gw[n_loc:]), pick(4, None if gb is None else gb[n_loc:]), None, None, None,
                None, None)


def _interleave32(w0, wf, wb):
    """[W_0; W_f; W_b] with rows interleaved in runs of 32 output features (the layout msda_zira_linear_16 expects)."""
    F, K = w0.shape
    return torch.stack([w0.view(F // 32, 32, K), wf.view(F // 32, 32, K), wb.view(F // 32, 32, K)], 1).reshape(3 * F, K).contiguous()


class ZiRaLinear16Function(Function):
    """Training-mode ZiRa projection beside a frozen linear, one tcgen05 GEMM (see msda_zira_linear_16):
    returns (y, zero_inter_loss).  Inputs: x [R, K] 16-bit; base (w0, b0); soft-frozen (wf, bf); branch (wb, bb);
    scaling s (1 element); optional row mask applied to y only."""

    @staticmethod
    def forward(ctx, x2d, row_mask, w0, b0, wf, bf, wb, bb, s):
        R, K = x2d.shape
        F = w0.shape[0]
        dt = x2d.dtype
        for name, t in (("w0", w0), ("b0", b0), ("wf", wf), ("bf", bf), ("wb", wb), ("bb", bb), ("scaling", s)):
            if t.dtype != dt:   # the GEMM reads raw 16-bit words: a promoted (fp32) stack would be read as garbage
                raise TypeError("ZiRaLinear16Function: %s is %s but the activation is %s -- cast the adapter weights with an "
                                "autograd-visible .to() at the call site" % (name, t.dtype, dt))
        w_stack = _interleave32(w0, wf, wb)
        bias3 = torch.cat([b0, bf, bb]).float()
        s32 = s.detach().float().reshape(1).contiguous()
        y = torch.empty((R, F), dtype=dt, device=x2d.device)
        pre = torch.empty((R, F), dtype=dt, device=x2d.device)
        adapter = torch.empty((R, F), dtype=dt, device=x2d.device)
        sums = torch.zeros(2, dtype=torch.float32, device=x2d.device)
        with torch.cuda.device(x2d.device):
            rc = _lib.lib().msda_zira_linear_16(x2d.data_ptr(), w_stack.data_ptr(), bias3.data_ptr(), s32.data_ptr(), R, K, F,
                                                y.data_ptr(), 0 if row_mask is None else row_mask.data_ptr(), pre.data_ptr(),
                                                adapter.data_ptr(), sums.data_ptr(), 1 if dt == torch.float16 else 0,
                                                _stream(x2d))
        _lib.check(rc, "msda_zira_linear_16")
        loss = (sums.sum() / (R * F)).to(dt)
        ctx.save_for_backward(x2d, row_mask, w0, wf, wb, s32, pre, adapter)
        return y, loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gloss):
        x2d, row_mask, w0, wf, wb, s32, pre, adapter = ctx.saved_tensors
        R, K = x2d.shape
        F = w0.shape[0]
        dt = x2d.dtype
        gl32 = gloss.detach().float().reshape(1).contiguous()
        stacked = torch.empty((R, 3 * F), dtype=dt, device=x2d.device)
        need = ctx.needs_input_grad
        scal = torch.zeros(1 + 3 * F, dtype=torch.float32, device=x2d.device)   # [d scaling | column sums of dY, dO, dB]
        ds, colsum = scal[:1], scal[1:]
        fused_colsum = (need[3] or need[5] or need[7]) and 256 % (F // 8) == 0 and 3 * F * 4 <= 48 * 1024
        with torch.cuda.device(x2d.device):
            rc = _lib.lib().msda_zira_bwd_prep_16(gy.contiguous().data_ptr(), pre.data_ptr(), adapter.data_ptr(),
                                                  0 if row_mask is None else row_mask.data_ptr(), s32.data_ptr(), gl32.data_ptr(),
                                                  R, F, stacked.data_ptr(), ds.data_ptr(),
                                                  colsum.data_ptr() if fused_colsum else 0, 1 if dt == torch.float16 else 0,
                                                  _stream(x2d))
        _lib.check(rc, "msda_zira_bwd_prep_16")
        s = s32.to(dt)
        gx = None
        if ctx.needs_input_grad[0]:   # dX = dY W_0 + dO W_f + dB (s W_b): one GEMM with K = 3F
            w_t = torch.cat([w0.t(), wf.t(), (wb * s).t()], 1).contiguous()
            # the tcgen05 GEMM holds Nout <= 2048; the im2col'd 3x3 level has K = 9*C_in (2304 / 6912): library product there
            gx = linear16(stacked, w_t) if K <= GEMM_MAX_N else stacked @ w_t.t()
        d_y, d_o, d_b = stacked[:, :F], stacked[:, F:2 * F], stacked[:, 2 * F:]
        if fused_colsum:
            bias_sum = lambda k, blk: colsum[k * F:(k + 1) * F]
        else:
            bias_sum = lambda k, blk: blk.float().sum(0)
        gw0 = d_y.t() @ x2d if need[2] else None
        gb0 = bias_sum(0, d_y).to(dt) if need[3] else None
        gwf = d_o.t() @ x2d if need[4] else None
        gbf = bias_sum(1, d_o).to(dt) if need[5] else None
        gwb = (d_b.t() @ x2d) * s if need[6] else None
        gbb = (bias_sum(2, d_b) * s32).to(dt) if need[7] else None
        gs = ds.to(dt) if need[8] else None
        return gx, None, gw0, gb0, gwf, gbf, gwb, gbb, gs


class QueryPostF32Function(Function):
    """fp32 modules: raw [R, 3*M*L*P] = [sampling_offsets | attention logits] pre-activations -> (sampling locations,
    softmax weights) in two kernels (forward) and one (backward) instead of ~10 eager ops (ms_deform_attn.py:290-319)."""

    @staticmethod
    def forward(ctx, raw, ref, spatial_shapes, M, L, P):
        R = raw.shape[0]
        raw = raw.contiguous()
        ref = ref.to(torch.float32).contiguous()
        loc = torch.empty((R, M, L, P, 2), dtype=torch.float32, device=raw.device)
        aw = torch.empty((R, M, L, P), dtype=torch.float32, device=raw.device)
        with torch.cuda.device(raw.device):
            rc = _lib.lib().msda_query_post_f32(raw.data_ptr(), ref.data_ptr(), ref.shape[-1], spatial_shapes.data_ptr(), R, M, L, P,
                                                loc.data_ptr(), aw.data_ptr(), _stream(raw))
        _lib.check(rc, "msda_query_post_f32")
        ctx.save_for_backward(ref, spatial_shapes, aw)
        ctx.dims = (M, L, P)
        return loc, aw

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loc, g_aw):
        ref, spatial_shapes, aw = ctx.saved_tensors
        M, L, P = ctx.dims
        R = aw.shape[0]
        n = 3 * M * L * P
        d_raw = torch.empty((R, n), dtype=torch.float32, device=aw.device)
        with torch.cuda.device(aw.device):
            rc = _lib.lib().msda_query_bwd_prep_16(g_loc.contiguous().data_ptr(), g_aw.contiguous().data_ptr(), aw.data_ptr(),
                                                   ref.data_ptr(), ref.shape[-1], spatial_shapes.data_ptr(), R, M, L, P,
                                                   d_raw.data_ptr(), n, 2, _stream(aw))
        _lib.check(rc, "msda_query_bwd_prep_16(fp32)")
        return d_raw, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# fp32 modules (the reference's default precision): the same fused structure on 3 x TF32 tcgen05 GEMMs
# (csrc/proj_gemm_f32.cu) -- fp32-GEMM accuracy (~1e-6 relative), so north_star's 1e-5 / 1e-4 bars hold.
# ---------------------------------------------------------------------------------------------------
def split_tf32(w):
    """nn.Linear weight [Nout, K] fp32 -> [2*Nout, K]: rows [0, Nout) = W with the 13 low mantissa bits cleared (what a
    TF32 operand keeps), rows [Nout, 2*Nout) = the exact remainder.  Done once per parameter version."""
    w = w.detach().float().contiguous()
    hi = (w.view(torch.int32) & -8192).view(torch.float32)
    return torch.cat([hi, w - hi], 0).contiguous()


def linear32(x2d, w_split, bias=None, row_mask=None, accum=None):
    """x2d [R, K] fp32 @ W^T (+ bias) (+ accum, in place) -> [R, Nout] fp32; rows with row_mask != 0 are zero."""
    R, K = x2d.shape
    Nout = w_split.shape[0] // 2
    assert x2d.is_contiguous() and x2d.dtype == torch.float32 and w_split.dtype == torch.float32 and w_split.shape[1] == K
    out = accum if accum is not None else torch.empty((R, Nout), dtype=torch.float32, device=x2d.device)
    if R == 0:
        return out
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_linear_f32(x2d.data_ptr(), w_split.data_ptr(), 0 if bias is None else bias.data_ptr(), R, K, Nout,
                                        0 if accum is None else accum.data_ptr(), out.data_ptr(), Nout,
                                        0 if row_mask is None else row_mask.data_ptr(), _stream(x2d))
    _lib.check(rc, "msda_linear_f32")
    return out


def query_proj32(q2d, w_cat_split, bias_cat, ref, ref_dim, spatial_shapes, M, L, P):
    R, K = q2d.shape
    loc = torch.empty((R, M, L, P, 2), dtype=torch.float32, device=q2d.device)
    aw = torch.empty((R, M, L, P), dtype=torch.float32, device=q2d.device)
    if R == 0:
        return loc, aw
    with torch.cuda.device(q2d.device):
        rc = _lib.lib().msda_query_proj_f32(q2d.data_ptr(), w_cat_split.data_ptr(), bias_cat.data_ptr(), ref.data_ptr(), ref_dim,
                                            spatial_shapes.data_ptr(), R, K, M, L, P, loc.data_ptr(), aw.data_ptr(), _stream(q2d))
    _lib.check(rc, "msda_query_proj_f32")
    return loc, aw


def supported32(embed_dim, M, L, P):
    lp = L * P
    return (embed_dim % 64 == 0 and embed_dim <= 768 and lp % 4 == 0 and 32 % lp == 0 and (3 * M * lp) % 64 == 0
            and 3 * M * lp <= 768 and L <= 16)


class Prepared32:
    """fp32 weights of one module, TF32-split ([w_hi; w_lo]) in the forward and transposed (dgrad) orientations."""

    def __init__(self, w_v, b_v, w_off, b_off, w_aw, b_aw, w_o, b_o):
        with torch.no_grad():
            w_cat = torch.cat([w_off.detach(), w_aw.detach()], 0)
            self.w_v, self.w_o, self.w_cat = split_tf32(w_v), split_tf32(w_o), split_tf32(w_cat)
            self.w_v_t, self.w_o_t, self.w_cat_t = split_tf32(w_v.detach().t()), split_tf32(w_o.detach().t()), split_tf32(w_cat.t())
            self.b_v, self.b_o = b_v.detach().float().contiguous(), b_o.detach().float().contiguous()
            self.b_cat = torch.cat([b_off.detach(), b_aw.detach()], 0).float().contiguous()


class FusedMSDeformAttnFunction32(Function):
    """Whole-module forward/backward on fp32 activations (batch-first, contiguous): 3 x TF32 GEMMs with the mask /
    sampling-location / softmax epilogues, the fp32 gather and scatter kernels, no library GEMM unless the module's own
    linears are trainable (their weight gradients)."""

    @staticmethod
    def forward(ctx, query, value_in, row_mask, reference_points, spatial_shapes, level_start_index, prep, M, L, P,
                im2col_step, w_v, b_v, w_off, b_off, w_aw, b_aw, w_o, b_o):
        N, Lq, C = query.shape
        S = value_in.shape[1]
        q2d, v2d = query.reshape(N * Lq, C), value_in.reshape(N * S, C)
        ref = reference_points.to(torch.float32).contiguous()
        ref_dim = ref.shape[-1]
        value = linear32(v2d, prep.w_v, prep.b_v, row_mask).view(N, S, M, C // M)
        loc, aw = query_proj32(q2d, prep.w_cat, prep.b_cat, ref, ref_dim, spatial_shapes, M, L, P)
        loc, aw = loc.view(N, Lq, M, L, P, 2), aw.view(N, Lq, M, L, P)
        core = _C.ms_deform_attn_forward(value, spatial_shapes, level_start_index, loc, aw, im2col_step)
        out = linear32(core.view(N * Lq, C), prep.w_o, prep.b_o).view(N, Lq, C)
        ctx.dims = (N, Lq, S, C, M, L, P, ref_dim, im2col_step)
        ctx.prep = prep
        ctx.wgrad = any(ctx.needs_input_grad[11:19])
        ctx.save_for_backward(query if ctx.wgrad else None, value_in if ctx.wgrad else None, row_mask, ref, spatial_shapes,
                              level_start_index, value, loc, aw, core if ctx.wgrad else None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        query, value_in, row_mask, ref, spatial_shapes, level_start_index, value, loc, aw, core = ctx.saved_tensors
        N, Lq, S, C, M, L, P, ref_dim, im2col_step = ctx.dims
        prep = ctx.prep
        g2d = grad_out.contiguous().view(N * Lq, C)
        d_core = linear32(g2d, prep.w_o_t)
        grad_value, grad_loc, grad_aw = _C.ms_deform_attn_backward(value, spatial_shapes, level_start_index, loc, aw,
                                                                   d_core.view(N, Lq, C), im2col_step)
        n_cat = 3 * M * L * P
        dq_cat = torch.empty((N * Lq, n_cat), dtype=torch.float32, device=value.device)
        with torch.cuda.device(value.device):
            rc = _lib.lib().msda_query_bwd_prep_16(grad_loc.data_ptr(), grad_aw.data_ptr(), aw.data_ptr(), ref.data_ptr(), ref_dim,
                                                   spatial_shapes.data_ptr(), N * Lq, M, L, P, dq_cat.data_ptr(), n_cat, 2,
                                                   _stream(value))
        _lib.check(rc, "msda_query_bwd_prep_16(fp32)")
        d_query = linear32(dq_cat, prep.w_cat_t).view(N, Lq, C) if ctx.needs_input_grad[0] else None
        gv2d = grad_value.view(N * S, C)
        # zeroing OUTPUT rows of this bias-free product == zeroing the masked rows of grad_value (backward of :287-288)
        d_value_in = linear32(gv2d, prep.w_v_t, None, row_mask).view(N, S, C) if ctx.needs_input_grad[1] else None
        grads = [None] * 8
        if ctx.wgrad:  # plain library GEMMs; the module's own linears are frozen in the ZiRa configuration
            n_loc = 2 * M * L * P
            q2d, v2d = query.reshape(N * Lq, C), value_in.reshape(N * S, C)
            gvm = gv2d if row_mask is None else gv2d.masked_fill(row_mask.bool()[:, None], 0.0)
            dw_cat = dq_cat.t() @ q2d
            db_cat = dq_cat.sum(0)
            grads = [gvm.t() @ v2d, gvm.sum(0), dw_cat[:n_loc], db_cat[:n_loc], dw_cat[n_loc:], db_cat[n_loc:],
                     g2d.t() @ core.view(N * Lq, C), g2d.sum(0)]
            grads = [g if ctx.needs_input_grad[11 + i] else None for i, g in enumerate(grads)]
        return (d_query, d_value_in, None, None, None, None, None, None, None, None, None, *grads)
