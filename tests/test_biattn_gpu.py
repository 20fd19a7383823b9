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
    e, es = rel_err(out.double(), ro), (stat.double() - rs).abs().max().item()
    assert e < 1e-2 and es < 2e-3, (e, es)


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
        e, es = rel_err(out.double(), ro), (stat.double() - rs).abs().max().item()
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
    out, _ = biattn.pv(a, b, x, H, scale, biattn._pad_mask(rm, B, LA, DEV), col_stat=col_stat, nsplit=nsplit)
    ro = _ref_given(a, b, x, H, scale, col_stat, rm)
    e = rel_err(out.double(), ro)
    assert e < 1e-2, e
    assert out[1, 5:40].abs().max().item() == 0.0
