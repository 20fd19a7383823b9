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
    assert torch.equal(ref_u, t("refpoint_embed_undetach")) and torch.equal(box, t("init_box_proposal"))
