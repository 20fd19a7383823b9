"""Row N4 of SURVEY.md 8(f): image <-> text fusion block.  The module is library attention (no CUDA extension needed), so
parity against the reference-generated fixture (tests/golden/make_golden.py: bi_attention_case, the reference's own
BiAttentionBlock in fp64) runs on CPU: outputs, input gradients, with and without masks, plus the oracle restatement that
materialises the attention matrix the way the reference does."""
import numpy as np
import pytest
import torch

from conftest import load_golden

from ziragroundingdino_b200.fuse_modules import BiAttentionBlock


def _load():
    g = load_golden("bi_attention")
    C, E, H = (int(x) for x in g["cfg"])
    blk = BiAttentionBlock(v_dim=C, l_dim=C, embed_dim=E, num_heads=H, dropout=0.0, drop_path=0.0).double()
    blk.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})   # strict: same keys
    return g, blk, (lambda k: torch.from_numpy(g[k]))


@pytest.mark.parametrize("use_sdpa", [False, True])
def test_bi_attention_block_matches_reference_fixture(use_sdpa, monkeypatch):
    from ziragroundingdino_b200.fuse_modules import BiMultiHeadAttention
    monkeypatch.setattr(BiMultiHeadAttention, "use_sdpa", use_sdpa)
    g, blk, t = _load()
    v, l = t("v").requires_grad_(True), t("l").requires_grad_(True)
    ov, ol = blk(v, l, attention_mask_v=t("mask_v"), attention_mask_l=t("mask_l"))
    assert np.abs(ov.detach().numpy() - g["out_v"]).max() < 1e-12 and np.abs(ol.detach().numpy() - g["out_l"]).max() < 1e-12
    ((ov * t("grad_out_v")).sum() + (ol * t("grad_out_l")).sum()).backward()
    assert np.abs(v.grad.numpy() - g["grad_v"]).max() < 1e-11 and np.abs(l.grad.numpy() - g["grad_l"]).max() < 1e-11
    ov, ol = blk(t("v"), t("l"))
    assert np.abs(ov.detach().numpy() - g["out_v_nomask"]).max() < 1e-12
    assert np.abs(ol.detach().numpy() - g["out_l_nomask"]).max() < 1e-12


def test_bi_attention_oracle_restatement_matches_reference():
    from oracle import cpu_encoder
    g, blk, t = _load()
    p = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    H = int(g["cfg"][2])
    ov, ol = cpu_encoder.bi_attention_block(p, t("v"), t("l"), t("mask_v"), t("mask_l"), H)
    assert np.abs(ov.numpy() - g["out_v"]).max() < 1e-12 and np.abs(ol.numpy() - g["out_l"]).max() < 1e-12
    ov, ol = cpu_encoder.bi_attention_block(p, t("v"), t("l"), None, None, H)
    assert np.abs(ov.numpy() - g["out_v_nomask"]).max() < 1e-12 and np.abs(ol.numpy() - g["out_l_nomask"]).max() < 1e-12


def test_bi_attention_extreme_logits_do_not_diverge_from_clamped_reference():
    """Logit spreads far beyond anything a trained model produces (thousands) but inside the reference's +-50000 clamp
    window: the unclamped attention still equals the reference's clamped arithmetic.  (Beyond the window the reference
    itself degenerates: a row lying entirely > 50000 below the global max is clamped flat and attends uniformly.)"""
    from oracle import cpu_encoder
    g, blk, t = _load()
    p = {k[6:]: torch.from_numpy(v).clone() for k, v in g.items() if k.startswith("param.")}
    p["attn.v_proj.weight"] *= 30.0
    p["attn.l_proj.weight"] *= 30.0
    blk.load_state_dict(p)
    H = int(g["cfg"][2])
    ov, ol = blk(t("v"), t("l"), attention_mask_v=t("mask_v"), attention_mask_l=t("mask_l"))
    rv, rl = cpu_encoder.bi_attention_block(p, t("v"), t("l"), t("mask_v"), t("mask_l"), H)
    assert torch.isfinite(ov).all() and torch.isfinite(ol).all()
    assert (ov - rv).abs().max() < 1e-9 and (ol - rl).abs().max() < 1e-9
