"""Bidirectional image <-> text attention core on the tcgen05 kernels of csrc/layer_biattn*.cu.

Host side: tensors stay in the layout the projections produce ([B, L, heads*256]; heads are addressed by the tensor
maps), masks and per-row statistics are padded to whole 128-row tiles, the split-column partial buffers are allocated
here, and ``BiAttentionCoreFunction`` ties the forward (two launches + a combine) to the backward (a row-dot, two
probability-transposed products, two logits-gradient products) -- reference fuse_modules.py:172-227 and autograd through it.
There is no fallback: a missing library raises (``_lib.lib()``).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .layer_ops import _stream

HD = 256          # head dimension the kernels are built for
TILE = 128


def supported(t, heads):
    return t.is_cuda and t.dtype in (torch.bfloat16, torch.float16) and t.shape[-1] == heads * HD


def _pad_len(L):
    return (L + TILE - 1) // TILE * TILE


def _pad_mask(mask, B, L, dev):
    """[B, L] bool (True = masked) or None -> uint8 [B, ceil(L/128)*128] with the padding marked masked."""
    out = torch.ones((B, _pad_len(L)), dtype=torch.uint8, device=dev)
    if mask is None:
        out[:, :L] = 0
    else:
        out[:, :L] = mask.to(torch.uint8)
    return out


def pad_stat(stat, fill):
    """[B, H, L] -> [B, H, ceil(L/128)*128] fp32 with `fill` in the padding (tests / callers with their own statistics)."""
    B, H, L = stat.shape
    out = torch.full((B, H, _pad_len(L)), fill, dtype=torch.float32, device=stat.device)
    out[:, :, :L] = stat
    return out


def default_splits(B, heads, LA, LB, dev, tile=TILE):
    """Column splits for the orientation with few stationary tiles: the split count whose work items (each a persistent
    CTA's share) finish in the fewest column tiles -- rounds of items over the SMs times tiles per item."""
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    units = B * heads * ((LA + TILE - 1) // TILE)
    ctiles = (LB + tile - 1) // tile
    if units >= sms:
        return 1
    best, best_cost = 1, None
    for ns in range(1, min(ctiles, (4 * sms) // units + 1) + 1):
        tps = (ctiles + ns - 1) // ns
        real = (ctiles + tps - 1) // tps
        cost = ((units * real + sms - 1) // sms) * tps + 0.02 * real      # small penalty per split: the combine pass
        if best_cost is None or cost < best_cost:
            best, best_cost = real, cost
    return best


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def pv(a, b, x, heads, scale, mask_padded, col_stat=None, nsplit=1, want_stat=True):
    """out[B, LA, E] = P . x with P from the logits scale * a b^T of each head.

    col_stat None : P = softmax over the LB axis (mask_padded = padded column mask); also returns the log2-domain
                    log-sum-exp per (b, head, a-row), padded to whole tiles with +inf, when want_stat.
    col_stat given: ([B, H, pad(LB)], +inf padding) P[i, j] = exp2(logit2[i, j] - col_stat[b, h, j]) for unmasked rows i
                    (mask_padded = padded row mask): the transposed probabilities of the other direction, not normalised.
    """
    B, LA, E = a.shape
    LB = b.shape[1]
    assert a.is_contiguous() and b.is_contiguous() and x.is_contiguous() and b.shape == x.shape and E == heads * HD
    L = _lib.lib()
    dev = a.device
    given = col_stat is not None
    assert not given or tuple(col_stat.shape) == (B, heads, _pad_len(LB))
    nsplit = L.msda_biattn_splits(LB, int(nsplit))
    out = torch.empty((B, LA, E), dtype=a.dtype, device=dev)
    stat = None
    if want_stat and not given:
        stat = torch.full((B, heads, _pad_len(LA)), float("inf"), dtype=torch.float32, device=dev)
    is_half = 1 if a.dtype == torch.float16 else 0
    po = pm = pl = None
    if nsplit > 1:
        items = B * heads * ((LA + TILE - 1) // TILE) * nsplit
        po = torch.empty((items, TILE, HD), dtype=torch.float32, device=dev)
        if not given:
            pm = torch.empty((items, TILE), dtype=torch.float32, device=dev)
            pl = torch.empty((items, TILE), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.msda_biattn_pv_16(a.data_ptr(), b.data_ptr(), x.data_ptr(), B, heads, LA, LB, float(scale), _ptr(mask_padded),
                                 _ptr(col_stat), out.data_ptr(), _ptr(stat) if nsplit == 1 else 0, _ptr(po), _ptr(pm), _ptr(pl),
                                 nsplit, is_half, _stream(a))
        _lib.check(rc, "msda_biattn_pv_16")
        if nsplit > 1:
            rc = L.msda_biattn_combine_16(_ptr(po), _ptr(pm), _ptr(pl), B, heads, LA, nsplit, 1 if given else 0, out.data_ptr(),
                                          _ptr(stat), is_half, _stream(a))
            _lib.check(rc, "msda_biattn_combine_16")
    return out, stat


def rowdot(d_o, o, heads):
    """delta[B, H, pad(L)] = sum over each head's 256 channels of d_o * o (zero padding)."""
    B, Lq, E = o.shape
    assert d_o.is_contiguous() and o.is_contiguous() and d_o.shape == o.shape and E == heads * HD
    lp = _pad_len(Lq)
    delta = torch.zeros((B, heads, lp), dtype=torch.float32, device=o.device)
    with torch.cuda.device(o.device):
        rc = _lib.lib().msda_biattn_rowdot_16(d_o.data_ptr(), o.data_ptr(), B, Lq, heads, lp, delta.data_ptr(),
                                              1 if o.dtype == torch.float16 else 0, _stream(o))
    _lib.check(rc, "msda_biattn_rowdot_16")
    return delta


def ds(a, d_oa, xa, b, xb, d_ob, heads, scale, mask_a_padded, mask_b_padded, lane_stat, lane_delta, col_stat, col_delta, nsplit=1,
       want_terms=False):
    """dA[B, LA, E] = scale * dS . b with dS the logits gradient of both softmax directions (csrc/layer_biattn_bwd.cu).
    want_terms: also return dS as a 16-bit [B, H, LA, ceil(LB/64)*64] tensor (unsplit launches only; the second pass adds
    onto the first with a TMA reduction).  `tn` turns it into the gradient of the streamed side without recomputing anything."""
    B, LA, E = a.shape
    LB = b.shape[1]
    for t in (a, d_oa, xa, b, xb, d_ob):
        assert t.is_contiguous() and t.dtype == a.dtype
    assert d_oa.shape == a.shape == xa.shape and xb.shape == b.shape == d_ob.shape and E == heads * HD
    assert tuple(lane_stat.shape) == tuple(lane_delta.shape) == (B, heads, _pad_len(LA))
    assert tuple(col_stat.shape) == tuple(col_delta.shape) == (B, heads, _pad_len(LB))
    L = _lib.lib()
    dev = a.device
    nsplit = L.msda_biattn_ds_splits(LB, int(nsplit))
    out = torch.empty((B, LA, E), dtype=a.dtype, device=dev)
    is_half = 1 if a.dtype == torch.float16 else 0
    if want_terms:
        assert nsplit == 1
        terms = torch.empty((B, heads, LA, (LB + 63) // 64 * 64), dtype=a.dtype, device=dev)
        with torch.cuda.device(dev):
            rc = L.msda_biattn_ds_terms_16(a.data_ptr(), d_oa.data_ptr(), xa.data_ptr(), b.data_ptr(), xb.data_ptr(), d_ob.data_ptr(), B,
                                           heads, LA, LB, float(scale), mask_a_padded.data_ptr(), mask_b_padded.data_ptr(),
                                           lane_stat.data_ptr(), lane_delta.data_ptr(), col_stat.data_ptr(), col_delta.data_ptr(),
                                           out.data_ptr(), terms.data_ptr(), is_half, _stream(a))
        _lib.check(rc, "msda_biattn_ds_terms_16")
        return out, terms
    po = None
    if nsplit > 1:
        po = torch.empty((B * heads * ((LA + TILE - 1) // TILE) * nsplit, TILE, HD), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.msda_biattn_ds_16(a.data_ptr(), d_oa.data_ptr(), xa.data_ptr(), b.data_ptr(), xb.data_ptr(), d_ob.data_ptr(), B, heads,
                                 LA, LB, float(scale), mask_a_padded.data_ptr(), mask_b_padded.data_ptr(), lane_stat.data_ptr(),
                                 lane_delta.data_ptr(), col_stat.data_ptr(), col_delta.data_ptr(), out.data_ptr(), _ptr(po), nsplit,
                                 is_half, _stream(a))
        _lib.check(rc, "msda_biattn_ds_16")
        if nsplit > 1:
            rc = L.msda_biattn_combine_16(_ptr(po), 0, 0, B, heads, LA, nsplit, 1, out.data_ptr(), 0, is_half, _stream(a))
            _lib.check(rc, "msda_biattn_combine_16")
    return out


def tn(terms, q, heads, scale, T, nsplit=None):
    """d k[B, T, E] = scale * dS^T . q per head (csrc/layer_biattn_tn.cu); terms = the dS of ds(..., want_terms=True)."""
    B, H, S, tpad = terms.shape
    assert terms.is_contiguous() and q.is_contiguous() and H == heads and tuple(q.shape) == (B, S, heads * HD) and tpad == (T + 63) // 64 * 64
    L = _lib.lib()
    dev = q.device
    if nsplit is None:      # work items cover two 128-token tiles each
        nsplit = default_splits(B, heads, (T + 1) // 2, S, dev, tile=64)
    nsplit = L.msda_biattn_tn_splits(S, int(nsplit))
    out = torch.empty((B, T, heads * HD), dtype=q.dtype, device=dev)
    is_half = 1 if q.dtype == torch.float16 else 0
    po = None
    if nsplit > 1:
        po = torch.empty((B * heads * ((T + TILE - 1) // TILE) * nsplit, TILE, HD), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.msda_biattn_tn_16(terms.data_ptr(), q.data_ptr(), B, heads, S, T, float(scale), out.data_ptr(), _ptr(po), nsplit, is_half,
                                 _stream(q))
        _lib.check(rc, "msda_biattn_tn_16")
        if nsplit > 1:
            rc = L.msda_biattn_combine_16(_ptr(po), 0, 0, B, heads, T, nsplit, 1, out.data_ptr(), 0, is_half, _stream(q))
            _lib.check(rc, "msda_biattn_combine_16")
    return out


# True (default): the rows-orientation logits-gradient launch stores its dS tiles (16 bit, B*H*S*256 elements) and d k is
# one product over them; False: d k recomputes logits and dP in the tokens orientation (nothing logits-sized in memory).
STORE_DS = __import__("os").environ.get("MSDA_B200_BIATTN_STORE_DS", "1") == "1"


def core_forward(q, k, val_v, val_l, mv, ml, heads, scale):
    """Both directions from padded byte masks: -> (out_v, out_l, stat_v, stat_l)."""
    B, S, _ = q.shape
    T = k.shape[1]
    out_v, stat_v = pv(q, k, val_l, heads, scale, ml)                                                  # rows orientation
    out_l, stat_l = pv(k, q, val_v, heads, scale, mv, nsplit=default_splits(B, heads, T, S, q.device))  # tokens orientation
    return out_v, out_l, stat_v, stat_l


def core_backward(q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml, d_out_v, d_out_l, heads, scale):
    """-> (d_q, d_k, d_val_v, d_val_l)."""
    B, S, _ = q.shape
    T = k.shape[1]
    dev = q.device
    delta_v, delta_l = rowdot(d_out_v, out_v, heads), rowdot(d_out_l, out_l, heads)
    ns_tok = default_splits(B, heads, T, S, dev)
    # values: the transposed probabilities of each direction times the other side's output gradient
    d_val_l, _ = pv(k, q, d_out_v, heads, scale, ml, col_stat=stat_v, nsplit=ns_tok)
    d_val_v, _ = pv(q, k, d_out_l, heads, scale, mv, col_stat=stat_l)
    # queries / keys: the logits gradient of both directions times the other operand
    if STORE_DS:
        d_q, terms = ds(q, d_out_v, val_v, k, val_l, d_out_l, heads, scale, mv, ml, stat_v, delta_v, stat_l, delta_l, want_terms=True)
        d_k = tn(terms, q, heads, scale, T)
    else:
        d_q = ds(q, d_out_v, val_v, k, val_l, d_out_l, heads, scale, mv, ml, stat_v, delta_v, stat_l, delta_l)
        d_k = ds(k, d_out_l, val_l, q, val_v, d_out_v, heads, scale, ml, mv, stat_l, delta_l, stat_v, delta_v,
                 nsplit=default_splits(B, heads, T, S, dev, tile=64))
    return d_q, d_k, d_val_v, d_val_l


class BiAttentionCoreFunction(Function):
    """(q [B,S,E], k [B,T,E], val_v [B,S,E], val_l [B,T,E]) -> (out_v [B,S,E], out_l [B,T,E]).

    out_v = softmax_T(scale q k^T, text mask) val_l ;  out_l = softmax_S(scale k q^T, image mask) val_v, per head of 256.
    """

    @staticmethod
    def forward(ctx, q, k, val_v, val_l, mask_v, mask_l, heads, scale):
        B, S, _ = q.shape
        T = k.shape[1]
        dev = q.device
        q, k, val_v, val_l = q.contiguous(), k.contiguous(), val_v.contiguous(), val_l.contiguous()
        mv, ml = _pad_mask(mask_v, B, S, dev), _pad_mask(mask_l, B, T, dev)
        out_v, out_l, stat_v, stat_l = core_forward(q, k, val_v, val_l, mv, ml, heads, scale)
        ctx.save_for_backward(q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml)
        ctx.heads, ctx.scale = heads, scale
        return out_v, out_l

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out_v, d_out_l):
        q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml = ctx.saved_tensors
        grads = core_backward(q, k, val_v, val_l, out_v, out_l, stat_v, stat_l, mv, ml, d_out_v.contiguous(), d_out_l.contiguous(),
                              ctx.heads, ctx.scale)
        return grads + (None, None, None, None)


def bi_attention_core(q, k, val_v, val_l, mask_v, mask_l, heads, scale):
    return BiAttentionCoreFunction.apply(q, k, val_v, val_l, mask_v, mask_l, heads, scale)
