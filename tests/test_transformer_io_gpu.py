"""Row N2 on the device: the pre/post steps of the encoder/decoder stacks (transformer_io.py) on CUDA tensors against
the reference-generated fixture tests/golden/transformer_io.npz -- same assertions as the CPU test, bit-exact where the
arithmetic is the same op sequence; plus the fused device kernels for the flatten + level-embedding step."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _g():
    g = load_golden("transformer_io")
    return g, (lambda k: torch.from_numpy(g[k]).to(DEV))


def test_flatten_levels_on_cuda_matches_reference():
    from ziragroundingdino_b200 import transformer_io as tio
    g, t = _g()
    L = g["shapes"].shape[0]
    srcs, masks, poss = [t("src%d" % i) for i in range(L)], [t("mask%d" % i) for i in range(L)], [t("pos%d" % i) for i in range(L)]
    src, mask, pos, shapes, sh, lsi, vr = tio.flatten_levels(srcs, masks, poss, t("level_embed"))
    assert src.is_cuda and sh.is_cuda and vr.is_cuda
    assert shapes == [tuple(int(v) for v in r) for r in g["shapes"]]
    assert torch.equal(src, t("src_flatten")) and torch.equal(mask, t("mask_flatten")) and torch.equal(pos, t("lvl_pos_embed_flatten"))
    assert torch.equal(sh, t("shapes")) and torch.equal(lsi, t("level_start_index")) and torch.equal(vr, t("valid_ratios"))
    pos_rows = torch.cat([p.flatten(2).transpose(1, 2) for p in poss], 1)
    assert torch.equal(tio.add_level_embed_rows(pos_rows, t("level_embed"), shapes), t("lvl_pos_embed_flatten"))


def test_proposals_and_topk_on_cuda_match_reference():
    from ziragroundingdino_b200 import transformer_io as tio
    g, t = _g()
    shapes = [tuple(int(v) for v in r) for r in g["shapes"]]
    om, op = tio.gen_encoder_output_proposals(t("memory"), t("mask_flatten"), shapes)
    # log() on the device may differ from the host libm in the last ulp: 1e-6 on finite entries, identical inf pattern
    want_p = t("output_proposals")
    assert torch.equal(torch.isinf(op), torch.isinf(want_p))
    fin = ~torch.isinf(want_p)
    assert (op[fin] - want_p[fin]).abs().max().item() < 1e-5
    assert torch.equal(om, t("output_memory"))
    om2, op2 = tio.gen_encoder_output_proposals(t("memory"), t("mask_flatten"), shapes, t("learnedwh"))
    assert torch.equal(torch.isinf(op2), torch.isinf(t("output_proposals_learnedwh")))
    assert torch.equal(om2, t("output_memory_learnedwh"))
    nq = g["topk_proposals"].shape[1]
    tgt, ref_u, box, idx = tio.select_topk_queries(t("output_memory"), t("class_logits"), t("coord_unselected"),
                                                   t("output_proposals"), nq)
    # top-k over distinct logits: same set in the same (sorted) order
    assert torch.equal(idx, t("topk_proposals")) and torch.equal(tgt, t("tgt_undetach"))
    assert torch.equal(ref_u, t("refpoint_embed_undetach"))
    # sigmoid of proposals that may differ from the CPU fixture in the last ulp of logf
    assert (box - t("init_box_proposal")).abs().max().item() < 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_device_kernels_match_eager_ops(dtype):
    """msda_flatten_levels / msda_level_valid_counts / msda_encoder_proposals vs the eager op sequence of the reference
    (transformer_for_adapter.py:238-262, utils.py:56-116) run with torch on the same device: bit-identical, including odd
    level sizes (tiles that run off the edge) and a channel count that is not a multiple of the 32-wide tile."""
    from ziragroundingdino_b200 import transformer_io as tio
    torch.manual_seed(4)
    N, C = 3, 80
    shapes = [(13, 21), (7, 11), (4, 6), (1, 3)]
    srcs = [torch.randn(N, C, h, w, device=DEV).to(dtype) for h, w in shapes]
    poss = [torch.randn(N, C, h, w, device=DEV).to(dtype) for h, w in shapes]
    masks = []
    for h, w in shapes:
        m = torch.zeros(N, h, w, dtype=torch.bool, device=DEV)
        m[1, h - h // 3:, :] = True
        m[2, :, w - w // 2:] = True
        masks.append(m)
    lvl = torch.randn(len(shapes), C, device=DEV).to(dtype)
    src, mask, pos, shp, sh, lsi, vr = tio.flatten_levels(srcs, masks, poss, lvl)
    want_src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    want_pos = torch.cat([p.flatten(2).transpose(1, 2) + lvl[i].view(1, 1, -1) for i, p in enumerate(poss)], 1)
    want_mask = torch.cat([m.flatten(1) for m in masks], 1)
    want_vr = torch.stack([tio.get_valid_ratio(m) for m in masks], 1)
    assert torch.equal(src, want_src) and torch.equal(pos, want_pos) and torch.equal(mask, want_mask) and torch.equal(vr, want_vr)
    assert shp == shapes and sh.tolist() == [list(s) for s in shapes]
    memory = torch.randn(N, src.shape[1], C, device=DEV).to(dtype)
    learnedwh = torch.tensor([0.3, -0.2], device=DEV)
    for lw in (None, learnedwh):
        om, op = tio.gen_encoder_output_proposals(memory, mask, shapes, lw)
        # eager restatement (the reference's op sequence) on the same device: requires_grad routes around the kernel
        om_e, op_e = tio.gen_encoder_output_proposals(memory.clone().requires_grad_(True), mask, shapes, lw)
        assert torch.equal(torch.isinf(op), torch.isinf(op_e))
        fin = ~torch.isinf(op_e)
        assert torch.equal(op[fin], op_e.detach()[fin])
        assert torch.equal(om, om_e.detach())
