"""The immediate caller of the hot path: the deformable transformer encoder layer.

Mirrors the reference's ``DeformableTransformerEncoderLayer`` (transformer_for_adapter.py:809-907; same
sub-module names, so checkpoint keys ``transformer.encoder.layers.{i}.self_attn.*`` / ``norm1`` /
``linear1`` ... load unchanged) with ``use_adapter=False`` as in the ZiRa configuration
(groundingdino/config/GroundingDINO_SwinT_OGC_rep.py:56).  The MSDeformAttn module inside is the
B200-native one; residual + LayerNorm are one fused kernel per direction and the FFN's two K = d_model products run on
the tcgen05 GEMM with ReLU / 1-bit-gate epilogues (layer_ops.py; the two K = d_ffn products stay library GEMMs); with
frozen parameters each half of the layer is ONE autograd Function whose dgrads accumulate in the GEMM epilogue
(blocks.py) -- SURVEY.md section 8(f) row N1.  ``DeformableEncoder`` is the 6-layer stack used by bench.py (BASELINE.json
config 2); the text-fusion and text-encoder sub-layers of the reference encoder loop
(transformer_for_adapter.py:563-612) are out of scope and omitted.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import blocks
from .layer_ops import add_layer_norm, ffn
from .ms_deform_attn import MultiScaleDeformableAttention


class DeformableTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 use_adapter=False, **unused):
        super().__init__()
        assert activation == "relu"
        if use_adapter:
            raise NotImplementedError("the bottleneck Adapter of the reference's encoder FFN is outside this path "
                                      "(the ZiRa configuration sets use_adapter=False)")
        self.self_attn = MultiScaleDeformableAttention(embed_dim=d_model, num_levels=n_levels, num_heads=n_heads,
                                                       num_points=n_points, batch_first=True)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = F.relu
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.use_adapter = False

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    # one autograd Function per half when everything inside is frozen (blocks.py): no elementwise gradient adds
    block_functions = True

    def forward_ffn(self, src):
        adapter_loss = src.new_zeros(1)
        out = blocks.ffn_block(self, src) if self.block_functions else None
        if out is not None:
            return out, adapter_loss
        src2 = ffn(src, self.linear1, self.linear2, self.dropout2.p, self.training)
        src = add_layer_norm(src, src2, self.norm2, self.dropout3.p, self.training)
        return src, adapter_loss

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None):
        out = None
        if self.block_functions:
            out = blocks.self_attn_block(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask)
        if out is not None:
            return self.forward_ffn(out)
        src2 = self.self_attn(query=self.with_pos_embed(src, pos), reference_points=reference_points, value=src,
                              spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                              key_padding_mask=key_padding_mask)
        src = add_layer_norm(src, src2, self.norm1, self.dropout1.p, self.training)
        return self.forward_ffn(src)


def get_reference_points(spatial_shapes, valid_ratios, device):
    """Pixel-centre grid per level, normalised by the valid extent (transformer_for_adapter.py:483-497).
    ``spatial_shapes``: sequence of (H, W) Python ints; ``valid_ratios``: [N, L, 2] (w, h)."""
    pts = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, dtype=torch.float32, device=device),
                                      torch.linspace(0.5, w - 0.5, w, dtype=torch.float32, device=device), indexing="ij")
        ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
        ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
        pts.append(torch.stack((ref_x, ref_y), -1))
    reference_points = torch.cat(pts, 1)
    return reference_points[:, :, None] * valid_ratios[:, None]


class DeformableEncoder(nn.Module):
    """``num_layers`` deformable encoder layers over flattened multi-scale features (config 2)."""

    def __init__(self, num_layers=6, d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_points=4, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([
            DeformableTransformerEncoderLayer(d_model, d_ffn, dropout, "relu", n_levels, n_heads, n_points)
            for _ in range(num_layers)])

    def forward(self, src, pos, spatial_shapes_host, spatial_shapes, level_start_index, valid_ratios, key_padding_mask):
        reference_points = get_reference_points(spatial_shapes_host, valid_ratios, src.device)
        out = src
        for layer in self.layers:
            out, _ = layer(out, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask)
        return out


def padded_batch_masks(shapes, N, device, generator=None, all_valid=False):
    """Synthetic key-padding masks of a batch padded to its largest image: per image a valid fraction of
    the padded canvas (short side 480-800 of 800, long side up to 1333 -- the reference's ODinW resize
    policy, groundingdino/config/configs/common/data/odinw/pothole.py:41-45).  Returns
    (mask [N, S] bool, valid_ratios [N, L, 2] (w, h))."""
    L = len(shapes)
    if all_valid:
        fh = torch.ones(N)
        fw = torch.ones(N)
    else:
        fh = torch.rand(N, generator=generator) * 0.4 + 0.6
        fw = torch.rand(N, generator=generator) * 0.4 + 0.6
        fh[0] = 1.0
        fw[-1] = 1.0          # the batch is padded to its largest member
    masks, ratios = [], torch.empty(N, L, 2)
    for lvl, (h, w) in enumerate(shapes):
        vh = torch.clamp((fh * h).ceil().long(), 1, h)
        vw = torch.clamp((fw * w).ceil().long(), 1, w)
        ys = torch.arange(h)[None, :, None] >= vh[:, None, None]
        xs = torch.arange(w)[None, None, :] >= vw[:, None, None]
        masks.append((ys | xs).reshape(N, h * w))
        ratios[:, lvl, 0] = vw.float() / w
        ratios[:, lvl, 1] = vh.float() / h
    return torch.cat(masks, 1).to(device), ratios.to(device)
