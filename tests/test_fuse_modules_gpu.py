"""GPU: row N4 -- BiAttentionBlock at the model's dimensions (v_dim = l_dim = 256, embed 1024, 4 heads of 256) in bf16 on
the tcgen05 attention core (and the two library formulations), against the oracle restatement (materialised attention matrix, fp64, CPU) on the
same parameters.  Outputs are LayerNorm-normalised inputs plus a gamma-scaled delta: bar 3e-2 absolute, 2e-2 relative
(Frobenius) on the deltas' input gradients."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("core", ["kernel", "kernel-frozen", "matmul", "sdpa"])
@pytest.mark.parametrize("masked", [True, False])
def test_bi_attention_block_bf16_vs_oracle(masked, core, monkeypatch):
    from oracle import cpu_encoder
    from ziragroundingdino_b200 import _lib
    from ziragroundingdino_b200.fuse_modules import BiAttentionBlock, BiMultiHeadAttention
    frozen = core == "kernel-frozen"
    core = "kernel" if frozen else core
    monkeypatch.setattr(BiMultiHeadAttention, "use_kernel", core == "kernel")
    monkeypatch.setattr(BiMultiHeadAttention, "use_sdpa", core == "sdpa")
    launches0 = _lib.launch_count()
    torch.manual_seed(9)
    B, n_img, n_text, C, E, H = 2, 1500, 48, 256, 1024, 4
    blk = BiAttentionBlock(v_dim=C, l_dim=C, embed_dim=E, num_heads=H, dropout=0.0, drop_path=0.0)
    with torch.no_grad():
        blk.gamma_v.fill_(0.5); blk.gamma_l.fill_(0.5)
    blk = blk.to(DEV).to(torch.bfloat16)
    if frozen:      # the whole block as one autograd node (FrozenBiAttentionBlockFunction)
        for prm in blk.parameters():
            prm.requires_grad_(False)
    v = torch.randn(B, n_img, C, device=DEV).to(torch.bfloat16).requires_grad_(True)
    l = torch.randn(B, n_text, C, device=DEV).to(torch.bfloat16).requires_grad_(True)
    mv = ml = None
    if masked:
        mv = torch.zeros(B, n_img, dtype=torch.bool, device=DEV); mv[1, -300:] = True
        ml = torch.zeros(B, n_text, dtype=torch.bool, device=DEV); ml[0, -10:] = True
    ov, ol = blk(v, l, attention_mask_v=mv, attention_mask_l=ml)
    gv, gl = torch.randn_like(ov), torch.randn_like(ol)
    ((ov.float() * gv.float()).sum() + (ol.float() * gl.float()).sum()).backward()
    # kernel core: 2 products + combine forward; 2 row-dots, 2 value products (+ combine), 2 logits-gradient products (+ combine)
    assert (_lib.launch_count() - launches0 >= (13 if frozen else 9)) == (core == "kernel")
    d = lambda t: t.detach().double().cpu()
    p = {k: d(t) for k, t in blk.state_dict().items()}
    vd, ld = d(v).requires_grad_(True), d(l).requires_grad_(True)
    rv, rl = cpu_encoder.bi_attention_block(p, vd, ld, None if mv is None else mv.cpu(), None if ml is None else ml.cpu(), H)
    ((rv * d(gv)).sum() + (rl * d(gl)).sum()).backward()
    assert (d(ov) - rv.detach()).abs().max().item() < 3e-2 and (d(ol) - rl.detach()).abs().max().item() < 3e-2
    assert rel_err(d(v.grad), vd.grad) < 2e-2 and rel_err(d(l.grad), ld.grad) < 2e-2
