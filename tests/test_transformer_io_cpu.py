"""Row N2 of SURVEY.md 8(f): the host-side steps either side of the encoder/decoder stacks, against a fixture generated
by running the reference's own code (tests/golden/make_golden.py: transformer_io_case).  Pure PyTorch host logic, so
these run on CPU; bit-exact where the arithmetic is the same sequence of ops (everything here)."""
import numpy as np
import torch

from conftest import load_golden

from ziragroundingdino_b200 import transformer_io as tio


def _g():
    g = load_golden("transformer_io")
    return g, (lambda k: torch.from_numpy(g[k]))


def test_flatten_levels_and_valid_ratios_match_reference():
    g, t = _g()
    L = g["shapes"].shape[0]
    srcs, masks, poss = [t("src%d" % i) for i in range(L)], [t("mask%d" % i) for i in range(L)], [t("pos%d" % i) for i in range(L)]
    src, mask, pos, shapes, sh, lsi, vr = tio.flatten_levels(srcs, masks, poss, t("level_embed"))
    assert shapes == [tuple(int(v) for v in r) for r in g["shapes"]]
    assert torch.equal(src, t("src_flatten")) and torch.equal(mask, t("mask_flatten")) and torch.equal(pos, t("lvl_pos_embed_flatten"))
    assert torch.equal(sh, t("shapes")) and torch.equal(lsi, t("level_start_index")) and torch.equal(vr, t("valid_ratios"))
    for i in range(L):
        assert torch.equal(tio.get_valid_ratio(masks[i]), t("valid_ratios")[:, i])
    # rows layout: one broadcast add of the per-token level embedding == the per-level adds
    pos_rows = torch.cat([p.flatten(2).transpose(1, 2) for p in poss], 1)
    assert torch.equal(tio.add_level_embed_rows(pos_rows, t("level_embed"), shapes), t("lvl_pos_embed_flatten"))


def test_gen_encoder_output_proposals_matches_reference():
    g, t = _g()
    shapes = [tuple(int(v) for v in r) for r in g["shapes"]]
    om, op = tio.gen_encoder_output_proposals(t("memory"), t("mask_flatten"), shapes)
    assert torch.equal(om, t("output_memory")) and torch.equal(op, t("output_proposals"))
    om, op = tio.gen_encoder_output_proposals(t("memory"), t("mask_flatten"), shapes, t("learnedwh"))
    assert torch.equal(om, t("output_memory_learnedwh")) and torch.equal(op, t("output_proposals_learnedwh"))
    assert torch.isinf(op[t("mask_flatten")]).all()


def test_select_topk_queries_matches_reference():
    g, t = _g()
    nq = g["topk_proposals"].shape[1]
    tgt, ref_u, box, idx = tio.select_topk_queries(t("output_memory"), t("class_logits"), t("coord_unselected"),
                                                   t("output_proposals"), nq)
    assert torch.equal(idx, t("topk_proposals")) and torch.equal(tgt, t("tgt_undetach"))
    assert torch.equal(ref_u, t("refpoint_embed_undetach")) and torch.equal(box, t("init_box_proposal"))
