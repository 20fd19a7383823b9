"""The step in front of the encoder: ZiRa-augmented input projection.

Mirrors what the reference builds and runs in ``GroundingDINO.__init__`` / ``forward``
(groundingdino_dual_zero_rep_branch.py:258-305 construction, :483-523 use): per feature level

    src_l = GroupNorm(32, hidden)( conv_0(x_l) + adapter_l(x_l) ),   loss += zero_inter_loss_l

with ``input_proj[l] = Sequential(Conv2d, GroupNorm)`` (1x1 for backbone levels, 3x3 / stride 2 / pad 1 for the
extra levels, the first of which reads the last backbone map and later ones the previous projected level) and
``input_proj_conv_adapter[l] = RepZeroConv2d`` of the same geometry.  Sub-module names equal the reference's, so the
checkpoint keys ``input_proj.{l}.{0,1}.*`` / ``input_proj_conv_adapter.{l}.*`` load unchanged and the reference's
``before_train`` name rule ("adapter") selects exactly the branch parameters.

B200 layout: features are taken channels-last as rows [N, H*W, C_in]; the projection is one fused tcgen05 GEMM per
level (base + soft-frozen + branch weights stacked, fold and SmoothL1 sums in the epilogue) writing [N, H*W, hidden]
rows, GroupNorm runs on those rows, and the levels are concatenated into the flattened [N, S, hidden] sequence the
encoder consumes (transformer_for_adapter.py:258-262) -- no NCHW round trip.  ``forward`` keeps the reference's NCHW
in / NCHW out contract for drop-in use.
"""
import torch
import torch.nn as nn

from .layer_ops import group_norm_rows
from .zira import RepZeroConv2d


class ZiRaInputProj(nn.Module):
    def __init__(self, num_channels=(192, 384, 768), hidden_dim=256, num_feature_levels=4, use_project_adapter=True,
                 norm_groups=32):
        super().__init__()
        self.num_feature_levels = num_feature_levels
        self.use_project_adapter = use_project_adapter
        num_channels = list(num_channels)
        nb = len(num_channels)
        if num_feature_levels == 1:
            num_channels, nb = num_channels[-1:], 1
        geom = [(c, 1, 1, 0) for c in num_channels]
        cin = num_channels[-1]
        for _ in range(num_feature_levels - nb):
            geom.append((cin, 3, 2, 1))
            cin = hidden_dim
        self.num_backbone_outs = nb
        self.input_proj = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, hidden_dim, kernel_size=k, stride=s, padding=p), nn.GroupNorm(norm_groups, hidden_dim))
            for c, k, s, p in geom])
        for proj in self.input_proj:                       # groundingdino_dual_zero_rep_branch.py:393-396
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        if use_project_adapter:
            self.input_proj_conv_adapter = nn.ModuleList([
                RepZeroConv2d(c, hidden_dim, kernel_size=k, stride=s, padding=p) for c, k, s, p in geom])

    def _level_rows(self, l, x_rows, hw):
        """One level on channels-last rows: returns (src rows [N, H'*W', hidden], (H', W'), loss or None)."""
        conv, norm = self.input_proj[l][0], self.input_proj[l][1]
        if self.use_project_adapter and self.input_proj_conv_adapter[l].rows_foldable(conv):
            y, out_hw, loss = self.input_proj_conv_adapter[l].forward_folded_rows(x_rows, hw, conv)
            if loss is None and not self.training:
                loss = torch.zeros(1).to(x_rows)           # what the reference's eval branch returns (:95)
            return group_norm_rows(y, norm), out_hw, loss
        N, HW, Cin = x_rows.shape
        x = x_rows.reshape(N, hw[0], hw[1], Cin).permute(0, 3, 1, 2)
        y = conv(x)
        loss = None
        if self.use_project_adapter:
            a, loss = self.input_proj_conv_adapter[l](x)
            y = y + a
        out_hw = tuple(y.shape[-2:])
        return group_norm_rows(y.flatten(2).transpose(1, 2), norm), out_hw, loss

    def forward_rows(self, feats_rows, feat_hw):
        """feats_rows[l] [N, H_l*W_l, C_l] channels-last backbone maps with feat_hw[l] = (H_l, W_l).
        Returns (src_flatten [N, S, hidden], spatial_shapes list[(H, W)] over all num_feature_levels,
        loss_conv_adapter or None)."""
        srcs, shapes, total = [], [], None
        for l in range(self.num_feature_levels):
            if l < self.num_backbone_outs:
                x_rows, hw = feats_rows[l], feat_hw[l]
            elif l == self.num_backbone_outs:
                x_rows, hw = feats_rows[-1], feat_hw[-1]   # :506-509: first extra level reads the last backbone map
            else:
                x_rows, hw = srcs[-1], shapes[-1]          # :516-519: later ones read the previous projected level
            y, out_hw, loss = self._level_rows(l, x_rows, hw)
            srcs.append(y)
            shapes.append(tuple(out_hw))
            if loss is not None:
                total = loss if total is None else total + loss
        return torch.cat(srcs, 1), shapes, total

    def forward(self, features):
        """Reference contract: list of NCHW backbone maps -> (list of NCHW projected maps, loss_conv_adapter)."""
        rows = [f.flatten(2).transpose(1, 2) for f in features]
        hw = [tuple(f.shape[-2:]) for f in features]
        flat, shapes, loss = self.forward_rows(rows, hw)
        outs, start = [], 0
        for h, w in shapes:
            outs.append(flat[:, start:start + h * w].transpose(1, 2).reshape(flat.shape[0], -1, h, w))
            start += h * w
        return outs, loss
