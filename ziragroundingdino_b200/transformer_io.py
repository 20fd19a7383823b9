"""The step either side of the encoder/decoder stacks: what the reference's
``Transformer.forward`` does before the encoder and between encoder and decoder
(transformer_for_adapter.py:228-262 flatten / level embedding / valid ratios; :300-340 two-stage proposals and top-k
query selection; utils.py:56-116 ``gen_encoder_output_proposals``).

B200-first in three ways: level shapes stay **on the host** as Python ints (the reference iterates a CUDA
``spatial_shapes`` tensor, which costs a device->host sync per level per call, utils.py:72); the flattened ``[N, S, C]``
layout produced by ``ZiRaInputProj.forward_rows`` is consumed as is (no per-level ``flatten(2).transpose(1, 2)`` copies);
and on CUDA tensors the two bulk steps are single kernels (csrc/layer_io.cu): ``flatten_levels`` = one tiled
NCHW -> [N, S, C] transposition of all levels with the level embedding added on the way (instead of ~6 eager kernels per
level), ``gen_encoder_output_proposals`` = one pass over the encoder memory (instead of ~40 eager kernels and five
[N, S, 4] temporaries).  CPU tensors take the eager PyTorch restatement (host plumbing, used by the CPU fixture test).
"""
import ctypes

import torch

from . import _lib

_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _device_path(*tensors):
    return all(t.is_cuda for t in tensors) and tensors[0].dtype in _DT and all(t.dtype == tensors[0].dtype for t in tensors)


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
    if (_device_path(*srcs, *pos_embeds) and (level_embed is None or (level_embed.is_cuda and level_embed.dtype == srcs[0].dtype))
            and all(m.is_cuda and m.dtype == torch.bool for m in masks) and not any(t.requires_grad for t in (*srcs, *pos_embeds))):
        return _flatten_levels_device(srcs, masks, pos_embeds, level_embed, shapes)
    src_flatten = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    mask_flatten = torch.cat([m.flatten(1) for m in masks], 1)
    pos = []
    for lvl, p in enumerate(pos_embeds):
        p = p.flatten(2).transpose(1, 2)
        pos.append(p if level_embed is None else p + level_embed[lvl].view(1, 1, -1))
    spatial_shapes, level_start_index = level_tensors(shapes, src_flatten.device)
    valid_ratios = torch.stack([get_valid_ratio(m) for m in masks], 1)
    return src_flatten, mask_flatten, torch.cat(pos, 1), shapes, spatial_shapes, level_start_index, valid_ratios


def _flatten_levels_device(srcs, masks, pos_embeds, level_embed, shapes):
    """flatten_levels on CUDA tensors: msda_flatten_levels + msda_level_valid_counts (no autograd: inputs are detached
    backbone features / position embeddings; level_embed is frozen in the ZiRa configuration)."""
    N, C = srcs[0].shape[:2]
    L = len(srcs)
    dev, dt = srcs[0].device, srcs[0].dtype
    S = sum(h * w for h, w in shapes)
    srcs = [t.contiguous() for t in srcs]
    poss = [t.contiguous() for t in pos_embeds]
    mks = [m.contiguous().view(torch.uint8) for m in masks]
    src_flatten = torch.empty((N, S, C), dtype=dt, device=dev)
    pos_flatten = torch.empty((N, S, C), dtype=dt, device=dev)
    mask_u8 = torch.empty((N, S), dtype=torch.uint8, device=dev)
    arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])
    hw = (ctypes.c_int * L)(*[h * w for h, w in shapes])
    le = None if level_embed is None else level_embed.detach().contiguous()
    lib = _lib.lib()
    with torch.cuda.device(dev):
        rc = lib.msda_flatten_levels(arr(srcs), arr(poss), arr(mks), hw, L, N, C, 0 if le is None else le.data_ptr(), _DT[dt],
                                     src_flatten.data_ptr(), pos_flatten.data_ptr(), mask_u8.data_ptr(), _stream(src_flatten))
        _lib.check(rc, "msda_flatten_levels")
        spatial_shapes, level_start_index = level_tensors(shapes, dev)
        counts = torch.empty((N, L, 2), dtype=torch.int32, device=dev)
        rc = lib.msda_level_valid_counts(mask_u8.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), N, S, L,
                                         counts.data_ptr(), _stream(src_flatten))
        _lib.check(rc, "msda_level_valid_counts")
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=dev)
    valid_ratios = counts.float() / wh[None]
    return src_flatten, mask_u8.view(torch.bool), pos_flatten, shapes, spatial_shapes, level_start_index, valid_ratios


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
    if (memory.is_cuda and memory.dtype in _DT and memory_padding_mask.is_cuda and memory_padding_mask.dtype == torch.bool
            and not memory.requires_grad and (memory.shape[-1] * memory.element_size()) % 16 == 0):
        return _proposals_device(memory, memory_padding_mask, shapes, learnedwh)
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


def _proposals_device(memory, memory_padding_mask, shapes, learnedwh):
    """gen_encoder_output_proposals on CUDA tensors: msda_level_valid_counts + msda_encoder_proposals (inference /
    detached use; with a memory that requires grad the eager path keeps autograd -- output_memory is a masked copy)."""
    N, S, C = memory.shape
    L = len(shapes)
    dev = memory.device
    memory = memory.contiguous()
    mask_u8 = memory_padding_mask.contiguous().view(torch.uint8)
    spatial_shapes, level_start_index = level_tensors(shapes, dev)
    counts = torch.empty((N, L, 2), dtype=torch.int32, device=dev)
    wh_base = (learnedwh.detach().float().sigmoid().reshape(2).contiguous() if learnedwh is not None
               else torch.full((2,), 0.05, dtype=torch.float32, device=dev))
    out_mem = torch.empty_like(memory)
    out_prop = torch.empty((N, S, 4), dtype=torch.float32, device=dev)
    lib = _lib.lib()
    with torch.cuda.device(dev):
        rc = lib.msda_level_valid_counts(mask_u8.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), N, S, L,
                                         counts.data_ptr(), _stream(memory))
        _lib.check(rc, "msda_level_valid_counts")
        rc = lib.msda_encoder_proposals(memory.data_ptr(), mask_u8.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                        counts.data_ptr(), wh_base.data_ptr(), N, S, L, C * memory.element_size(),
                                        out_mem.data_ptr(), out_prop.data_ptr(), _stream(memory))
        _lib.check(rc, "msda_encoder_proposals")
    return out_mem, out_prop


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
