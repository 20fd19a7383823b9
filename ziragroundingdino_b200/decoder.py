"""The caller on the other side of the hot path: the deformable transformer decoder layer.

Mirrors the reference's ``DeformableTransformerDecoderLayer`` (transformer_for_adapter.py:910-1073; same constructor
arguments and sub-module names, so checkpoint keys ``transformer.decoder.layers.{i}.cross_attn.*`` / ``ca_text`` /
``self_attn`` / ``norm{1,2,3}`` / ``linear{1,2}`` load unchanged) with ``use_adapter=False`` as in the ZiRa
configuration (GroundingDINO_SwinT_OGC_rep.py:56).  The image cross-attention is the B200-native
MultiScaleDeformableAttention (900 queries over the encoder memory); residual + LayerNorm and the FFN use the fused
kernels of layer_ops.py on 16-bit CUDA tensors; query self-attention and text cross-attention (900 x 900 and
900 x n_text dense attention) stay ``nn.MultiheadAttention`` library calls -- they are not on the path SURVEY.md 8 names.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .layer_ops import add_layer_norm, ffn
from .ms_deform_attn import MultiScaleDeformableAttention


def _p(drop):
    return getattr(drop, "p", 0.0)


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 use_text_feat_guide=False, use_text_cross_attention=False, use_adapter=False, **unused):
        super().__init__()
        assert activation == "relu" and not use_text_feat_guide
        if use_adapter:
            raise NotImplementedError("the bottleneck Adapter of the reference's decoder FFN is outside this path "
                                      "(the ZiRa configuration sets use_adapter=False)")
        mk_drop = lambda: nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.cross_attn = MultiScaleDeformableAttention(embed_dim=d_model, num_levels=n_levels, num_heads=n_heads,
                                                        num_points=n_points, batch_first=True)
        self.dropout1 = mk_drop()
        self.norm1 = nn.LayerNorm(d_model)
        if use_text_cross_attention:
            self.ca_text = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
            self.catext_dropout = mk_drop()
            self.catext_norm = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = mk_drop()
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = F.relu
        self.dropout3 = mk_drop()
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = mk_drop()
        self.norm3 = nn.LayerNorm(d_model)
        self.key_aware_proj = None
        self.use_text_feat_guide = use_text_feat_guide
        self.use_text_cross_attention = use_text_cross_attention
        self.use_adapter = False

    def rm_self_attn_modules(self):
        self.self_attn = None
        self.dropout2 = None
        self.norm2 = None

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, tgt):
        adapter_loss = tgt.new_zeros(1)
        tgt2 = ffn(tgt, self.linear1, self.linear2, _p(self.dropout3), self.training)
        tgt = add_layer_norm(tgt, tgt2, self.norm3, _p(self.dropout4), self.training)
        return tgt, adapter_loss

    def forward(self, tgt, tgt_query_pos=None, tgt_query_sine_embed=None, tgt_key_padding_mask=None,
                tgt_reference_points=None, memory_text=None, text_attention_mask=None, memory=None,
                memory_key_padding_mask=None, memory_level_start_index=None, memory_spatial_shapes=None, memory_pos=None,
                self_attn_mask=None, cross_attn_mask=None):
        """tgt / tgt_query_pos [nq, bs, C]; tgt_reference_points [nq, bs, L, 4]; memory [S, bs, C] (seq-first, as the
        reference passes them, transformer_for_adapter.py:1024-1073).  Returns (tgt, adapter_loss)."""
        assert cross_attn_mask is None
        if self.self_attn is not None:
            q = k = self.with_pos_embed(tgt, tgt_query_pos)
            # need_weights=False: the reference discards the head-averaged attention map it asks for ([0] of the pair,
            # transformer_for_adapter.py:1044); not asking lets the library run fused scaled-dot-product attention instead of
            # materialising 900 x 900 probabilities per head (bmm + softmax + mean: ~0.6 ms per layer at 8 images)
            tgt2 = self.self_attn(q, k, tgt, attn_mask=self_attn_mask, need_weights=False)[0]
            tgt = add_layer_norm(tgt, tgt2, self.norm2, _p(self.dropout2), self.training)
        if self.use_text_cross_attention:
            tgt2 = self.ca_text(self.with_pos_embed(tgt, tgt_query_pos), memory_text.transpose(0, 1),
                                memory_text.transpose(0, 1), key_padding_mask=text_attention_mask, need_weights=False)[0]
            tgt = add_layer_norm(tgt, tgt2, self.catext_norm, _p(self.catext_dropout), self.training)
        tgt2 = self.cross_attn(query=self.with_pos_embed(tgt, tgt_query_pos).transpose(0, 1),
                               reference_points=tgt_reference_points.transpose(0, 1).contiguous(),
                               value=memory.transpose(0, 1), spatial_shapes=memory_spatial_shapes,
                               level_start_index=memory_level_start_index,
                               key_padding_mask=memory_key_padding_mask).transpose(0, 1)
        tgt = add_layer_norm(tgt, tgt2, self.norm1, _p(self.dropout1), self.training)
        return self.forward_ffn(tgt)
