"""Image <-> text feature fusion of the encoder loop.

Mirrors the reference's ``BiMultiHeadAttention`` / ``BiAttentionBlock`` (fuse_modules.py:99-254, :258-307; used once per
encoder layer, transformer_for_adapter.py:578-593) -- same constructor arguments, parameter names and outputs -- but
never materialises the [heads, S, n_text] attention matrix the reference builds twice in fp32 (182 MB per image at
S = 22 223, 256 text tokens): both directions are ordinary attention problems,

    image <- text :  softmax_text ( q k^T )           v_l         q = v_proj(v) * scale, k = l_proj(l)
    text <- image :  softmax_image( k q^T )           v_v         (the transposed logits)

because every step the reference applies to the logits before a softmax is a per-row shift (global max :174, row max
:187) or a clamp at +-50 000 (:177-195), and the clamp only binds on entries whose probability is already
exp(-50000) = 0 -- as long as each row's own maximum lies within 50 000 of the global maximum, i.e. for any logits a
trained model produces (they are O(10); tested up to spreads of thousands).  Outside that window the reference itself
degenerates (a row entirely below the window is clamped flat and attends uniformly); that quirk is not reproduced.  So each
direction is a plain softmax attention: computed from logits kept in the activation dtype (default), or as
two fused scaled-dot-product-attention calls that never materialise it (``use_sdpa``).

Default for 16-bit CUDA tensors with 256-wide heads and no attention dropout (the reference configs: fusion_dropout = 0.0):
the attention core runs on this package's own tcgen05 kernels (``biattn.py`` / csrc/layer_biattn*.cu) -- both directions
from the projections' natural [B, L, heads*256] layout, logits and probabilities never leaving the SM, backward by
recomputation from per-row log-sum-exp statistics.  The six projections are plain GEMMs and stay on the library.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, biattn
from .layer_ops import _stream, derived


class DropPath(nn.Module):
    """Stochastic depth per sample (the reference takes timm's; fuse_modules.py:11)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


class BiMultiHeadAttention(nn.Module):
    # False (default): materialised logits in the activation dtype, innermost-axis softmaxes, library GEMMs (fastest
    # measured at the model's sizes, see tools/bench_fusion.py).  True: two fused scaled-dot-product-attention calls -- nothing of size n_img x n_text is ever
    # materialised (memory-lean; its backward for the 256-query x 22 223-key direction parallelises poorly today).
    use_sdpa = False
    # True (default): the tcgen05 attention core when the tensors qualify (see the module docstring); the paths above otherwise
    use_kernel = os.environ.get("MSDA_B200_BIATTN", "1") == "1"

    def __init__(self, v_dim, l_dim, embed_dim, num_heads, dropout=0.1, cfg=None):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.v_dim, self.l_dim = v_dim, l_dim
        assert self.head_dim * num_heads == embed_dim, \
            f"embed_dim must be divisible by num_heads (got `embed_dim`: {embed_dim} and `num_heads`: {num_heads})."
        self.scale = self.head_dim ** (-0.5)
        self.dropout = dropout
        self.v_proj = nn.Linear(v_dim, embed_dim)
        self.l_proj = nn.Linear(l_dim, embed_dim)
        self.values_v_proj = nn.Linear(v_dim, embed_dim)
        self.values_l_proj = nn.Linear(l_dim, embed_dim)
        self.out_v_proj = nn.Linear(embed_dim, v_dim)
        self.out_l_proj = nn.Linear(embed_dim, l_dim)
        self.stable_softmax_2d = True
        self.clamp_min_for_underflow = True
        self.clamp_max_for_overflow = True
        self._reset_parameters()

    def _reset_parameters(self):
        for lin in (self.v_proj, self.l_proj, self.values_v_proj, self.values_l_proj, self.out_v_proj, self.out_l_proj):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.fill_(0)

    def _heads(self, t):
        b, n, _ = t.shape
        # head-major and contiguous ([B, H, n, d]): strided head views push the library's batched GEMMs onto slow
        # unaligned kernels (measured: 3.4 ms of a 10 ms block at the model's sizes)
        return t.view(b, n, self.num_heads, self.head_dim).transpose(1, 2).contiguous()

    def forward(self, v, l, attention_mask_v=None, attention_mask_l=None):
        """v [bs, n_img, v_dim], l [bs, n_text, l_dim]; masks [bs, n_img] / [bs, n_text] bool, True = padding.
        Returns (attn_output_v [bs, n_img, v_dim], attn_output_l [bs, n_text, l_dim])."""
        bsz, n_img, _ = v.shape
        n_text = l.shape[1]
        n_img_in = n_img
        p_drop = self.dropout if self.training else 0.0
        if self.use_kernel and v.is_cuda and self.head_dim == biattn.HD and p_drop == 0.0:
            q = self.v_proj(v)                      # the 1/sqrt(head_dim) scale is applied to the logits inside the kernel
            if q.dtype in (torch.bfloat16, torch.float16):
                k, val_v, val_l = self.l_proj(l), self.values_v_proj(v), self.values_l_proj(l)
                if k.dtype == q.dtype == val_v.dtype == val_l.dtype:
                    out_v, out_l = biattn.bi_attention_core(q, k, val_v, val_l, attention_mask_v, attention_mask_l,
                                                            self.num_heads, self.scale)
                    return self.out_v_proj(out_v), self.out_l_proj(out_l)
        if not self.use_sdpa and n_img % 8 and v.is_cuda:
            # S = 22 223 is odd: a [.., n_text, S] logits matrix then has rows that are not 16-byte aligned and the library
            # falls back to its unaligned GEMM / softmax kernels.  Pad the image tokens to a multiple of 8 with masked rows.
            pad = 8 - n_img % 8
            v = F.pad(v, (0, 0, 0, pad))
            if attention_mask_v is None:
                attention_mask_v = v.new_zeros((bsz, n_img), dtype=torch.bool)
            attention_mask_v = F.pad(attention_mask_v, (0, pad), value=True)
            n_img += pad
        q = self._heads(self.v_proj(v) * self.scale)
        k = self._heads(self.l_proj(l))
        val_v = self._heads(self.values_v_proj(v))
        val_l = self._heads(self.values_l_proj(l))
        p = self.dropout if self.training else 0.0
        if self.use_sdpa:
            keep_l = None if attention_mask_l is None else ~attention_mask_l[:, None, None, :]
            keep_v = None if attention_mask_v is None else ~attention_mask_v[:, None, None, :]
            out_v = F.scaled_dot_product_attention(q, k, val_l, attn_mask=keep_l, dropout_p=p, scale=1.0)     # image <- text
            out_l = F.scaled_dot_product_attention(k, q, val_v, attn_mask=keep_v, dropout_p=p, scale=1.0)     # text <- image
        else:
            # Both logits matrices in the activation dtype ([B, H, n_img, n_text] and its transpose, 182 MB each in bf16 at
            # N=4, S=22 223, 256 tokens; the reference holds them in fp32 plus temporaries).  The transpose is a second
            # tensor-core product rather than a strided softmax: torch's softmax over a non-innermost axis of this shape is
            # ~100x slower than over the innermost one (measured, tools/bench_fusion.py).
            w_v = torch.matmul(q, k.transpose(-1, -2))                              # image rows, text columns
            w_l = torch.matmul(k, q.transpose(-1, -2))                              # text rows, image columns
            if attention_mask_l is not None:      # in place: the products' backward does not need their outputs
                w_v.masked_fill_(attention_mask_l[:, None, None, :], float("-inf"))
            if attention_mask_v is not None:
                w_l.masked_fill_(attention_mask_v[:, None, None, :], float("-inf"))
            out_v = torch.matmul(F.dropout(torch.softmax(w_v, dim=-1), p, self.training), val_l)
            out_l = torch.matmul(F.dropout(torch.softmax(w_l, dim=-1), p, self.training), val_v)
        out_v = out_v.transpose(1, 2).reshape(bsz, n_img, self.embed_dim)
        out_l = out_l.transpose(1, 2).reshape(bsz, n_text, self.embed_dim)
        return self.out_v_proj(out_v[:, :n_img_in]), self.out_l_proj(out_l)


def _layer_norm16(x2, norm, shift=None):
    """LayerNorm on the package's kernel: -> (y, y + shift or None, mean, rstd)."""
    R, C = x2.shape
    y = torch.empty_like(x2)
    y2 = torch.empty_like(x2) if shift is not None else None
    mean = torch.empty(R, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(R, dtype=torch.float32, device=x2.device)
    g32, b32 = derived(norm.weight, "f32"), derived(norm.bias, "f32")
    with torch.cuda.device(x2.device):
        rc = _lib.lib().msda_layernorm_fwd_16(x2.data_ptr(), g32.data_ptr(), b32.data_ptr(), R, C, float(norm.eps), y.data_ptr(),
                                              0 if y2 is None else y2.data_ptr(), 0 if shift is None else shift.data_ptr(),
                                              mean.data_ptr(), rstd.data_ptr(), 1 if x2.dtype == torch.float16 else 0, _stream(x2))
    _lib.check(rc, "msda_layernorm_fwd_16")
    return y, y2, mean, rstd, g32


def _layer_norm16_bwd(dy2, x2, g32, mean, rstd):
    dx = torch.empty_like(x2)
    R, C = x2.shape
    with torch.cuda.device(x2.device):
        rc = _lib.lib().msda_add_layernorm_bwd_16(dy2.data_ptr(), x2.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), R, C,
                                                  dx.data_ptr(), 1 if x2.dtype == torch.float16 else 0, _stream(x2))
    _lib.check(rc, "msda_add_layernorm_bwd_16")
    return dx


class FrozenBiAttentionBlockFunction(Function):
    """The whole BiAttentionBlock (reference fuse_modules.py:258-307) as ONE autograd node when none of its parameters is
    trained (the ZiRa configuration: gradients only flow THROUGH the fusion layers) and dropout / drop-path are inactive:
    LayerNorms on the package's kernel (the image-side one also emits `LN(v) + gamma_v * b_out`), gamma folded into the
    output projections so that `v + gamma * (x W^T + b)` is one GEMM with a residual operand, and in the backward the three
    gradients reaching LN(v) -- residual, q projection, value projection -- are accumulated by the dgrad GEMMs themselves.
    Removes 2 multiplies, 4 adds and 2 slow LayerNorm kernels per direction from the eager composition."""

    @staticmethod
    def forward(ctx, v, l, mask_v, mask_l, blk):
        a = blk.attn
        H, E = a.num_heads, a.embed_dim
        B, S, C = v.shape
        T = l.shape[1]
        dev = v.device
        v2, l2 = v.reshape(B * S, C).contiguous(), l.reshape(B * T, l.shape[-1]).contiguous()
        w_ov = (blk.gamma_v[:, None] * a.out_v_proj.weight).contiguous()            # [v_dim, E]
        w_ol = (blk.gamma_l[:, None] * a.out_l_proj.weight).contiguous()
        shift_v = blk.gamma_v.float() * a.out_v_proj.bias.float()
        shift_l = blk.gamma_l.float() * a.out_l_proj.bias.float()
        v_ln, v_res, mean_v, rstd_v, g_v = _layer_norm16(v2, blk.layer_norm_v, shift_v)
        l_ln, l_res, mean_l, rstd_l, g_l = _layer_norm16(l2, blk.layer_norm_l, shift_l)
        q = F.linear(v_ln, a.v_proj.weight, a.v_proj.bias).view(B, S, E)           # unscaled: the kernel scales the logits
        val_v = F.linear(v_ln, a.values_v_proj.weight, a.values_v_proj.bias).view(B, S, E)
        k = F.linear(l_ln, a.l_proj.weight, a.l_proj.bias).view(B, T, E)
        val_l = F.linear(l_ln, a.values_l_proj.weight, a.values_l_proj.bias).view(B, T, E)
        mv, ml = biattn._pad_mask(mask_v, B, S, dev), biattn._pad_mask(mask_l, B, T, dev)
        out_v, out_l, stat_v, stat_l = biattn.core_forward(q, k, val_v, val_l, mv, ml, H, a.scale)
        v_out = torch.addmm(v_res, out_v.view(B * S, E), w_ov.t())
        l_out = torch.addmm(l_res, out_l.view(B * T, E), w_ol.t())
        ctx.save_for_backward(v2, l2, q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml, mean_v, rstd_v, mean_l, rstd_l,
                              g_v, g_l, w_ov, w_ol)
        ctx.blk = blk
        ctx.shapes = (v.shape, l.shape)
        return v_out.view(v.shape), l_out.view(l.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_v_out, g_l_out):
        (v2, l2, q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml, mean_v, rstd_v, mean_l, rstd_l, g_v, g_l, w_ov,
         w_ol) = ctx.saved_tensors
        a = ctx.blk.attn
        B, S, E = q.shape
        T = k.shape[1]
        gv2, gl2 = g_v_out.reshape(v2.shape).contiguous(), g_l_out.reshape(l2.shape).contiguous()
        d_out_v = torch.mm(gv2, w_ov).view(B, S, E)
        d_out_l = torch.mm(gl2, w_ol).view(B, T, E)
        d_q, d_k, d_val_v, d_val_l = biattn.core_backward(q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml, d_out_v, d_out_l,
                                                          a.num_heads, a.scale)
        d_vln = torch.addmm(gv2, d_q.view(B * S, E), a.v_proj.weight)
        d_vln.addmm_(d_val_v.view(B * S, E), a.values_v_proj.weight)
        d_lln = torch.addmm(gl2, d_k.view(B * T, E), a.l_proj.weight)
        d_lln.addmm_(d_val_l.view(B * T, E), a.values_l_proj.weight)
        d_v = _layer_norm16_bwd(d_vln, v2, g_v, mean_v, rstd_v).view(ctx.shapes[0])
        d_l = _layer_norm16_bwd(d_lln, l2, g_l, mean_l, rstd_l).view(ctx.shapes[1])
        return d_v, d_l, None, None, None


class BiAttentionBlock(nn.Module):
    def __init__(self, v_dim, l_dim, embed_dim, num_heads, dropout=0.1, drop_path=0.0, init_values=1e-4, cfg=None):
        super().__init__()
        self.layer_norm_v = nn.LayerNorm(v_dim)
        self.layer_norm_l = nn.LayerNorm(l_dim)
        self.attn = BiMultiHeadAttention(v_dim=v_dim, l_dim=l_dim, embed_dim=embed_dim, num_heads=num_heads, dropout=dropout)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.gamma_v = nn.Parameter(init_values * torch.ones((v_dim)), requires_grad=True)
        self.gamma_l = nn.Parameter(init_values * torch.ones((l_dim)), requires_grad=True)

    def _fused_ok(self, v, l):
        a = self.attn
        dp = self.drop_path
        return (a.use_kernel and v.is_cuda and v.dtype in (torch.bfloat16, torch.float16) and l.dtype == v.dtype
                and a.head_dim == biattn.HD and not (self.training and a.dropout > 0.0)
                and (isinstance(dp, nn.Identity) or not self.training or dp.drop_prob == 0.0)
                and not torch.is_autocast_enabled() and v.shape[-1] % 8 == 0 and l.shape[-1] % 8 == 0
                and all(p.dtype == v.dtype and not p.requires_grad for p in self.parameters()))

    def forward(self, v, l, attention_mask_v=None, attention_mask_l=None):
        if self._fused_ok(v, l):
            return FrozenBiAttentionBlockFunction.apply(v, l, attention_mask_v, attention_mask_l, self)
        v = self.layer_norm_v(v)
        l = self.layer_norm_l(l)
        delta_v, delta_l = self.attn(v, l, attention_mask_v=attention_mask_v, attention_mask_l=attention_mask_l)
        v = v + self.drop_path(self.gamma_v * delta_v)
        l = l + self.drop_path(self.gamma_l * delta_l)
        return v, l
