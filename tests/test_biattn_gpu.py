"""GPU: the tcgen05 bidirectional-attention kernels (csrc/layer_biattn.cu, row N4) against an fp64 torch restatement of
softmax attention on the same 16-bit inputs.  Bars: 1e-2 relative (Frobenius) on 16-bit outputs -- the probabilities are
rounded to 16 bit before the second product, as in any fused attention kernel; statistics 2e-3 absolute (log2 domain)."""
import math

import pytest
import torch


pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOG2E = 1.4426950408889634


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _heads(t, H):
    B, L, E = t.shape
    return t.double().view(B, L, H, E // H).transpose(1, 2)


def _logits2(a, b, H, scale):
    return (_heads(a, H) @ _heads(b, H).transpose(-1, -2)) * (scale * LOG2E)       # [B, H, LA, LB], log2 domain


def _ref_online(a, b, x, H, scale, col_mask):
    s2 = _logits2(a, b, H, scale)
    if col_mask is not None:
        s2 = s2.masked_fill(col_mask[:, None, None, :], float("-inf"))
    lse2 = torch.logsumexp(s2 * math.log(2.0), dim=-1) / math.log(2.0)
    p = torch.exp2(s2 - lse2[..., None])
    o = p @ _heads(x, H)
    B, _, LA, hd = o.shape
    return o.transpose(1, 2).reshape(B, LA, H * hd), lse2


def _ref_given(a, b, x, H, scale, col_stat, row_mask):
    s2 = _logits2(a, b, H, scale)
    p = torch.exp2(s2 - col_stat.double()[:, :, None, :])
    if row_mask is not None:
        p = p.masked_fill(row_mask[:, None, :, None], 0.0)
    o = p @ _heads(x, H)
    B, _, LA, hd = o.shape
    return o.transpose(1, 2).reshape(B, LA, H * hd)


def _mk(B, L, H, dtype, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(B, L, H * 256, device=DEV, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("LA,LB,nsplit,masked", [(300, 200, 1, True), (128, 256, 1, False), (200, 1000, 3, True),
                                                 (256, 1500, 12, False), (77, 50, 1, True)])
def test_pv_online_matches_softmax_attention(LA, LB, nsplit, masked, dtype):
    from ziragroundingdino_b200 import biattn
    B, H, scale = 2, 2, 1.0 / 16
    a, b, x = _mk(B, LA, H, dtype, seed=1), _mk(B, LB, H, dtype, seed=2), _mk(B, LB, H, dtype, seed=3)
    cm = None
    if masked:
        cm = torch.zeros(B, LB, dtype=torch.bool, device=DEV)
        cm[0, LB // 3: LB // 3 + 17] = True
        cm[1, -9:] = True
    out, stat = biattn.pv(a, b, x, H, scale, biattn._pad_mask(cm, B, LB, DEV), nsplit=nsplit)
    ro, rs = _ref_online(a, b, x, H, scale, cm)
    e = rel_err(out.double(), ro)
    es = (stat[:, :, :LA].double() - rs).abs().max().item()
    assert e < 1e-2 and es < 2e-3, (e, es)
    assert torch.isinf(stat[:, :, LA:]).all()


def test_pv_online_rescales_when_later_tiles_dominate():
    """Logits that grow along the streamed axis force the running reference to be raised (accumulator rescaled in TMEM)."""
    from ziragroundingdino_b200 import biattn
    B, H, LA, LB, scale = 1, 2, 130, 900, 1.0 / 16
    a, x = _mk(B, LA, H, torch.bfloat16, seed=4), _mk(B, LB, H, torch.bfloat16, seed=5)
    grow = torch.linspace(0.2, 6.0, LB, device=DEV)[None, :, None]
    b = (_mk(B, LB, H, torch.float32, seed=6) * grow).to(torch.bfloat16)
    for nsplit in (1, 2):
        out, stat = biattn.pv(a, b, x, H, scale, biattn._pad_mask(None, B, LB, DEV), nsplit=nsplit)
        ro, rs = _ref_online(a, b, x, H, scale, None)
        e, es = rel_err(out.double(), ro), (stat[:, :, :LA].double() - rs).abs().max().item()
        assert e < 1e-2 and es < 5e-3, (nsplit, e, es)


@pytest.mark.parametrize("LA,LB,nsplit", [(300, 200, 1), (200, 1000, 4), (128, 128, 1)])
def test_pv_given_statistics_reproduces_transposed_probabilities(LA, LB, nsplit):
    from ziragroundingdino_b200 import biattn
    B, H, scale, dtype = 2, 2, 1.0 / 16, torch.bfloat16
    a, b, x = _mk(B, LA, H, dtype, seed=7), _mk(B, LB, H, dtype, seed=8), _mk(B, LB, H, dtype, seed=9)
    rm = torch.zeros(B, LA, dtype=torch.bool, device=DEV)
    rm[1, 5:40] = True
    # statistics of the other direction: softmax over the rows (axis LA) of each column, masked rows excluded
    s2 = _logits2(a, b, H, scale).masked_fill(rm[:, None, :, None], float("-inf"))
    col_stat = (torch.logsumexp(s2 * math.log(2.0), dim=2) / math.log(2.0)).float().contiguous()      # [B, H, LB]
    out, _ = biattn.pv(a, b, x, H, scale, biattn._pad_mask(rm, B, LA, DEV), col_stat=biattn.pad_stat(col_stat, float("inf")), nsplit=nsplit)
    ro = _ref_given(a, b, x, H, scale, col_stat, rm)
    e = rel_err(out.double(), ro)
    assert e < 1e-2, e
    assert out[1, 5:40].abs().max().item() == 0.0


def _ref_core(q, k, vv, vl, mv, ml, H, scale):
    """fp64 autograd restatement of both directions (reference fuse_modules.py:172-227 without the no-op clamps)."""
    s = (_heads(q, H) @ _heads(k, H).transpose(-1, -2)) * scale                       # [B, H, S, T]
    sv = s if ml is None else s.masked_fill(ml[:, None, None, :], float("-inf"))
    sl = s.transpose(-1, -2)
    sl = sl if mv is None else sl.masked_fill(mv[:, None, None, :], float("-inf"))
    ov = torch.softmax(sv, -1) @ _heads(vl, H)
    ol = torch.softmax(sl, -1) @ _heads(vv, H)
    B, _, S, hd = ov.shape
    return ov.transpose(1, 2).reshape(B, S, H * hd), ol.transpose(1, 2).reshape(B, ol.shape[2], H * hd)


@pytest.mark.parametrize("store_ds", [True, False])
@pytest.mark.parametrize("S,T,masked", [(300, 200, True), (1000, 256, False), (2500, 48, True)])
def test_core_function_forward_backward_vs_fp64(S, T, masked, store_ds, monkeypatch):
    from ziragroundingdino_b200 import biattn
    monkeypatch.setattr(biattn, "STORE_DS", store_ds)
    B, H, scale, dtype = 2, 2, 1.0 / 16, torch.bfloat16
    q, k = _mk(B, S, H, dtype, seed=11).requires_grad_(True), _mk(B, T, H, dtype, seed=12).requires_grad_(True)
    vv, vl = _mk(B, S, H, dtype, seed=13).requires_grad_(True), _mk(B, T, H, dtype, seed=14).requires_grad_(True)
    mv = ml = None
    if masked:
        mv = torch.zeros(B, S, dtype=torch.bool, device=DEV); mv[1, -S // 5:] = True
        ml = torch.zeros(B, T, dtype=torch.bool, device=DEV); ml[0, -7:] = True
    ov, ol = biattn.bi_attention_core(q, k, vv, vl, mv, ml, H, scale)
    gv, gl = _mk(B, S, H, dtype, seed=15), _mk(B, T, H, dtype, seed=16)
    torch.autograd.backward([ov, ol], [gv, gl])
    ref_in = [t.detach().double().requires_grad_(True) for t in (q, k, vv, vl)]
    rv, rl = _ref_core(*ref_in, mv, ml, H, scale)
    torch.autograd.backward([rv, rl], [gv.double(), gl.double()])
    errs = {"out_v": rel_err(ov.double(), rv.detach()), "out_l": rel_err(ol.double(), rl.detach())}
    for name, t, r in zip(("d_q", "d_k", "d_val_v", "d_val_l"), (q, k, vv, vl), ref_in):
        errs[name] = rel_err(t.grad.double(), r.grad)
    assert all(e < 1e-2 for e in errs.values()), errs


@pytest.mark.parametrize("B,H,S,T,dtype", [(1, 1, 50, 7, torch.bfloat16), (1, 3, 700, 300, torch.bfloat16), (2, 4, 129, 64, torch.float16),
                                           (3, 2, 1024, 513, torch.bfloat16)])
def test_core_function_edge_shapes(B, H, S, T, dtype):
    """Fewer image tokens than one tile, more than 256 text tokens (several token tiles / groups), one head, fp16."""
    from ziragroundingdino_b200 import biattn
    scale = 1.0 / 16
    q, k = _mk(B, S, H, dtype, seed=21).requires_grad_(True), _mk(B, T, H, dtype, seed=22).requires_grad_(True)
    vv, vl = _mk(B, S, H, dtype, seed=23).requires_grad_(True), _mk(B, T, H, dtype, seed=24).requires_grad_(True)
    mv = torch.zeros(B, S, dtype=torch.bool, device=DEV); mv[0, S // 2:] = True
    ml = torch.zeros(B, T, dtype=torch.bool, device=DEV); ml[-1, :T // 3] = True
    ov, ol = biattn.bi_attention_core(q, k, vv, vl, mv, ml, H, scale)
    gv, gl = _mk(B, S, H, dtype, seed=25), _mk(B, T, H, dtype, seed=26)
    torch.autograd.backward([ov, ol], [gv, gl])
    ref_in = [t.detach().double().requires_grad_(True) for t in (q, k, vv, vl)]
    rv, rl = _ref_core(*ref_in, mv, ml, H, scale)
    torch.autograd.backward([rv, rl], [gv.double(), gl.double()])
    errs = {"out_v": rel_err(ov.double(), rv.detach()), "out_l": rel_err(ol.double(), rl.detach())}
    for name, t, r in zip(("d_q", "d_k", "d_val_v", "d_val_l"), (q, k, vv, vl), ref_in):
        errs[name] = rel_err(t.grad.double(), r.grad)
    assert all(e < 1e-2 for e in errs.values()), errs


def test_block_under_inference_mode_and_noncontiguous_inputs():
    from ziragroundingdino_b200 import _lib
    from ziragroundingdino_b200.fuse_modules import BiAttentionBlock
    torch.manual_seed(3)
    blk = BiAttentionBlock(256, 256, 1024, 4, dropout=0.0, drop_path=0.0).to(DEV).bfloat16().eval()
    v = torch.randn(2, 2, 400, 256, device=DEV).bfloat16()[:, 0]            # a strided view
    l = torch.randn(2, 33, 256, device=DEV).bfloat16()
    n0 = _lib.launch_count()
    with torch.inference_mode():
        ov, ol = blk(v, l)
    assert _lib.launch_count() - n0 >= 3
    blk.attn.use_kernel = False
    try:
        with torch.no_grad():
            rv, rl = blk(v.contiguous(), l)
    finally:
        del blk.attn.use_kernel
    assert (ov.float() - rv.float()).abs().max().item() < 6e-2 and (ol.float() - rl.float()).abs().max().item() < 6e-2
