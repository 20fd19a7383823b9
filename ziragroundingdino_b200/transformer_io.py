"""The step either side of the encoder/decoder stacks: what the reference's
``Transformer.forward`` does before the encoder and between encoder and decoder
(transformer_for_adapter.py:228-262 flatten / level embedding / valid ratios; :300-340 two-stage proposals and top-k
query selection; utils.py:56-116 ``gen_encoder_output_proposals``).

Host-side PyTorch plumbing, B200-first in two ways: level shapes stay **on the host** as Python ints (the reference
iterates a CUDA ``spatial_shapes`` tensor, which costs a device->host sync per level per call, utils.py:72), and the
flattened ``[N, S, C]`` layout produced by ``ZiRaInputProj.forward_rows`` is consumed as is (no per-level
``flatten(2).transpose(1, 2)`` copies).
"""
import torch


def get_valid_ratio(mask):
    """mask [N, H, W] bool (True = padding) -> [N, 2] (w, h) valid fraction (transformer_for_adapter.py:216-223)."""
    _, H, W = mask.shape
    valid_h = torch.sum(~mask[:, :, 0], 1)
    valid_w = torch.sum(~mask[:, 0, :], 1)
    return torch.stack([valid_w.float() / W, valid_h.float() / H], -1)


def level_tensors(shapes, device):
    """Host shapes [(H, W), ...] -> (spatial_shapes [L, 2] int64, level_start_index [L] int64) on ``device``."""
    sh = torch.as_tensor(shapes, dtype=torch.long, device=device)
    return sh, torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))


def flatten_levels(srcs, masks, pos_embeds, level_embed=None):
    """Reference contract (transformer_for_adapter.py:238-262): lists of NCHW maps, [N, H, W] masks and NCHW position
    embeddings -> (src_flatten [N, S, C], mask_flatten [N, S], lvl_pos_embed_flatten [N, S, C], shapes (host list),
    spatial_shapes, level_start_index, valid_ratios [N, L, 2])."""
    shapes = [tuple(s.shape[-2:]) for s in srcs]
    src_flatten = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    mask_flatten = torch.cat([m.flatten(1) for m in masks], 1)
    pos = []
    for lvl, p in enumerate(pos_embeds):
        p = p.flatten(2).transpose(1, 2)
        pos.append(p if level_embed is None else p + level_embed[lvl].view(1, 1, -1))
    spatial_shapes, level_start_index = level_tensors(shapes, src_flatten.device)
    valid_ratios = torch.stack([get_valid_ratio(m) for m in masks], 1)
    return src_flatten, mask_flatten, torch.cat(pos, 1), shapes, spatial_shapes, level_start_index, valid_ratios


def add_level_embed_rows(pos_rows, level_embed, shapes):
    """Rows layout: pos_rows [N, S, C] already flattened; adds ``level_embed[l]`` to the tokens of level l in ONE
    broadcast add (the per-level ``pos_embed + level_embed[lvl]`` of :250-253)."""
    counts = torch.tensor([h * w for h, w in shapes], device=pos_rows.device)
    per_token = torch.repeat_interleave(level_embed, counts, dim=0, output_size=int(sum(h * w for h, w in shapes)))
    return pos_rows + per_token[None].to(pos_rows.dtype)


def gen_encoder_output_proposals(memory, memory_padding_mask, shapes, learnedwh=None):
    """utils.py:56-116 with host ``shapes`` [(H, W), ...]: (output_memory [N, S, C], output_proposals [N, S, 4] unsigmoid).
    Same arithmetic, one level at a time; no device->host synchronisation."""
    N = memory.shape[0]
    proposals, cur = [], 0
    for lvl, (H, W) in enumerate(shapes):
        m = memory_padding_mask[:, cur:cur + H * W].view(N, H, W)
        valid_h = torch.sum(~m[:, :, 0], 1)
        valid_w = torch.sum(~m[:, 0, :], 1)
        grid_y, grid_x = torch.meshgrid(torch.linspace(0, H - 1, H, dtype=torch.float32, device=memory.device),
                                        torch.linspace(0, W - 1, W, dtype=torch.float32, device=memory.device), indexing="ij")
        grid = torch.cat([grid_x.unsqueeze(-1), grid_y.unsqueeze(-1)], -1)
        scale = torch.cat([valid_w.unsqueeze(-1), valid_h.unsqueeze(-1)], 1).view(N, 1, 1, 2)
        grid = (grid.unsqueeze(0).expand(N, -1, -1, -1) + 0.5) / scale
        if learnedwh is not None:
            wh = torch.ones_like(grid) * learnedwh.sigmoid() * (2.0 ** lvl)
        else:
            wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
        proposals.append(torch.cat((grid, wh), -1).view(N, -1, 4))
        cur += H * W
    output_proposals = torch.cat(proposals, 1)
    valid = ((output_proposals > 0.01) & (output_proposals < 0.99)).all(-1, keepdim=True)
    output_proposals = torch.log(output_proposals / (1 - output_proposals))
    drop = memory_padding_mask.unsqueeze(-1) | ~valid
    output_proposals = output_proposals.masked_fill(drop, float("inf"))
    output_memory = memory.masked_fill(drop, float(0))
    return output_memory, output_proposals


def select_topk_queries(output_memory, class_logits, coord_unselected, output_proposals, num_queries):
    """Two-stage query selection (transformer_for_adapter.py:311-329): top-k tokens by their best class logit.
    Returns (tgt_undetach [N, nq, C], refpoint_embed_undetach [N, nq, 4] unsigmoid, init_box_proposal [N, nq, 4],
    topk_proposals [N, nq])."""
    topk_logits = class_logits.max(-1)[0]
    topk_proposals = torch.topk(topk_logits, num_queries, dim=1)[1]
    idx4 = topk_proposals.unsqueeze(-1).expand(-1, -1, 4)
    refpoint_embed_undetach = torch.gather(coord_unselected, 1, idx4)
    init_box_proposal = torch.gather(output_proposals, 1, idx4).sigmoid()
    tgt_undetach = torch.gather(output_memory, 1, topk_proposals.unsqueeze(-1).expand(-1, -1, output_memory.shape[-1]))
    return tgt_undetach, refpoint_embed_undetach, init_box_proposal, topk_proposals
