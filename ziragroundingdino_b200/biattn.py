"""Bidirectional image <-> text attention core on the tcgen05 kernels of csrc/layer_biattn.cu.

Host side of the two kernels: tensors stay in the layout the projections produce ([B, L, heads*256], heads addressed by
the tensor maps), masks are padded to whole 128-row tiles, the split-column partial buffers are allocated here.
There is no fallback: a missing library raises (``_lib.lib()``).
"""
import torch

from . import _lib
from .layer_ops import _stream

HD = 256          # head dimension the kernels are built for
TILE = 128


def supported(q, l_len, heads):
    return (q.is_cuda and q.dtype in (torch.bfloat16, torch.float16) and q.shape[-1] == heads * HD and l_len >= 1)


def _pad_mask(mask, B, L, dev):
    """[B, L] bool (True = masked) or None -> uint8 [B, ceil(L/128)*128] with the padding marked masked."""
    Lp = (L + TILE - 1) // TILE * TILE
    out = torch.ones((B, Lp), dtype=torch.uint8, device=dev)
    if mask is None:
        out[:, :L] = 0
    else:
        out[:, :L] = mask.to(torch.uint8)
    return out


def default_splits(B, heads, LA, LB, dev):
    """Column splits for the orientation with few stationary tiles: fill the SMs about twice."""
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    units = B * heads * ((LA + TILE - 1) // TILE)
    if units >= sms:
        return 1
    return max(1, min((LB + TILE - 1) // TILE, (2 * sms + units - 1) // units))


def pv(a, b, x, heads, scale, mask_padded, col_stat=None, nsplit=1, want_stat=True):
    """out[B, LA, E] = P . x with P from the logits scale * a b^T of each head.

    col_stat None : P = softmax over the LB axis (mask_padded = padded column mask); also returns the log2-domain
                    log-sum-exp per (b, head, a-row) when want_stat.
    col_stat given: P[i, j] = exp2(logit2[i, j] - col_stat[b, h, j]) for unmasked rows i (mask_padded = padded row mask),
                    i.e. the transposed probabilities of the other direction; nothing is normalised.
    """
    B, LA, E = a.shape
    LB = b.shape[1]
    assert a.is_contiguous() and b.is_contiguous() and x.is_contiguous() and b.shape == x.shape and E == heads * HD
    L = _lib.lib()
    dev = a.device
    given = col_stat is not None
    nsplit = L.msda_biattn_splits(LB, int(nsplit))
    out = torch.empty((B, LA, E), dtype=a.dtype, device=dev)
    stat = torch.empty((B, heads, LA), dtype=torch.float32, device=dev) if (want_stat and not given) else None
    is_half = 1 if a.dtype == torch.float16 else 0
    po = pm = pl = None
    if nsplit > 1:
        items = B * heads * ((LA + TILE - 1) // TILE) * nsplit
        po = torch.empty((items, TILE, HD), dtype=torch.float32, device=dev)
        if not given:
            pm = torch.empty((items, TILE), dtype=torch.float32, device=dev)
            pl = torch.empty((items, TILE), dtype=torch.float32, device=dev)
    ptr = lambda t: 0 if t is None else t.data_ptr()
    with torch.cuda.device(dev):
        rc = L.msda_biattn_pv_16(a.data_ptr(), b.data_ptr(), x.data_ptr(), B, heads, LA, LB, float(scale), ptr(mask_padded),
                                 ptr(col_stat), out.data_ptr(), ptr(stat) if nsplit == 1 else 0, ptr(po), ptr(pm), ptr(pl),
                                 nsplit, is_half, _stream(a))
        _lib.check(rc, "msda_biattn_pv_16")
        if nsplit > 1:
            rc = L.msda_biattn_combine_16(ptr(po), ptr(pm), ptr(pl), B, heads, LA, nsplit, 1 if given else 0, out.data_ptr(),
                                          ptr(stat), is_half, _stream(a))
            _lib.check(rc, "msda_biattn_combine_16")
    return out, stat
