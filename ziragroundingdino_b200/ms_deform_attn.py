"""B200-native drop-in for the reference's ``groundingdino/models/GroundingDINO/ms_deform_attn.py``.

Public names, constructor, attributes, ``forward`` signature, ``state_dict`` keys and error
behaviour mirror the reference (ms_deform_attn.py:38-87 for the autograd Function, :133-355 for the
module) so a checkpoint and a call site written for the reference work unchanged; the arithmetic
runs in the sm_100a kernels behind ``ziragroundingdino_b200._C`` / include/msda_b200.h.

Differences, all deliberate:
  * there is NO CPU path.  The reference falls back to ``multi_scale_deformable_attn_pytorch`` for
    CPU tensors (ms_deform_attn.py:345-348); here a CPU tensor raises ``RuntimeError``.
  * bf16 is supported (the reference op dispatches float/double only, ms_deform_attn_cuda.cu:65);
    fp16 and bf16 run natively with 16-bit value/output, fp32 locations/weights and fp32
    accumulation, which is what the reference's fp16 up-cast computes (ms_deform_attn.py:326-344).
  * the ``sum(H*W) == num_value`` check (ms_deform_attn.py:284) costs the reference a device->host
    sync per call; here the host copy of ``spatial_shapes`` is cached per tensor version.
  * ZiRa: ``add_zira_branches()`` attaches re-parameterisable side branches with the semantics of the
    reference's ``RepZeroLinear`` (groundingdino_dual_zero_rep_branch.py:105-135) to ``value_proj``
    and ``output_proj`` (see zira.py).
"""
import math
import warnings
import weakref
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.init import constant_, xavier_uniform_

from . import _C, fused
from .zira import RepZeroLinear


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MultiScaleDeformableAttnFunction(Function):
    """Same call signature as the reference Function (ms_deform_attn.py:38-87)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = _C.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                           sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, loc, aw = ctx.saved_tensors
        grad_value, grad_loc, grad_aw = _C.ms_deform_attn_backward(
            value, shapes, level_start, loc, aw, grad_output.contiguous(), ctx.im2col_step)
        if grad_value.dtype != value.dtype:
            grad_value = grad_value.to(value.dtype)
        return grad_value, None, None, grad_loc, grad_aw, None


def multi_scale_deformable_attn_pytorch(*args, **kwargs):
    """The reference's CPU implementation (ms_deform_attn.py:90-130) has no counterpart here."""
    raise RuntimeError("ziragroundingdino_b200 has no CPU / PyTorch fallback for multi-scale deformable "
                       "attention; use MultiScaleDeformableAttnFunction on CUDA tensors")


_SHAPE_CACHE = {}


def _version_of(t):
    """Version counter of a tensor, or a fresh object for inference tensors (which have none): never equal, so no caching."""
    return object() if t.is_inference() else t._version


def _host_shapes(spatial_shapes):
    """Host copy of ``spatial_shapes`` as a tuple of (H, W).  Cached per tensor *object* (weak reference) and
    version counter, so a call costs one device->host sync the first time a given tensor is seen and none
    afterwards (the reference syncs on every call, ms_deform_attn.py:284)."""
    if not spatial_shapes.is_cuda or spatial_shapes.is_inference():
        # inference tensors carry no version counter (reading ``_version`` raises): no cache, one sync, as the reference
        return tuple((int(h), int(w)) for h, w in spatial_shapes.tolist())
    key = id(spatial_shapes)
    hit = _SHAPE_CACHE.get(key)
    if hit is not None and hit[0]() is spatial_shapes and hit[1] == spatial_shapes._version:
        return hit[2]
    if len(_SHAPE_CACHE) > 64:
        for k in [k for k, v in _SHAPE_CACHE.items() if v[0]() is None]:
            del _SHAPE_CACHE[k]
    val = tuple((int(h), int(w)) for h, w in spatial_shapes.tolist())
    _SHAPE_CACHE[key] = (weakref.ref(spatial_shapes), spatial_shapes._version, val)
    return val


class MultiScaleDeformableAttention(nn.Module):
    """Multi-Scale Deformable Attention (Deformable-DETR), reference ms_deform_attn.py:133-355.

    Args:
        embed_dim (int): embedding dimension. Default: 256.
        num_heads (int): attention heads. Default: 8.
        num_levels (int): feature levels. Default: 4.
        num_points (int): sampling points per head per level. Default: 4.
        img2col_step (int): kept for API compatibility (only its divisibility check is observable).
        batch_first (bool): ``(bs, n, embed_dim)`` if True else ``(n, bs, embed_dim)``. Default: False.
    """

    def __init__(self, embed_dim: int = 256, num_heads: int = 8, num_levels: int = 4, num_points: int = 4,
                 img2col_step: int = 64, batch_first: bool = False):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise ValueError("embed_dim must be divisible by num_heads, but got {} and {}".format(
                embed_dim, num_heads))
        head_dim = embed_dim // num_heads
        self.batch_first = batch_first
        if not _is_power_of_2(head_dim):
            warnings.warn(
                """
                You'd better set d_model in MSDeformAttn to make sure that
                each dim of the attention head a power of 2, which is more efficient.
                """
            )
        self.im2col_step = img2col_step
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.num_levels = num_levels
        self.num_points = num_points
        self.sampling_offsets = nn.Linear(embed_dim, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dim, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dim, embed_dim)
        self.output_proj = nn.Linear(embed_dim, embed_dim)
        # ZiRa side branches (None until add_zira_branches()); not part of the reference state_dict
        self.value_proj_adapter = None
        self.output_proj_adapter = None
        self.zero_inter_loss = None  # set by forward() in training mode when branches exist
        self.init_weights()

    def _reset_parameters(self):
        return self.init_weights()

    def init_weights(self):
        """Reference initialisation (ms_deform_attn.py:194-217)."""
        constant_(self.sampling_offsets.weight.data, 0.0)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2)
        grid_init = grid_init.repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.0)
        constant_(self.attention_weights.bias.data, 0.0)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.0)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.0)

    def freeze_sampling_offsets(self):
        print("Freeze sampling offsets")
        self.sampling_offsets.weight.requires_grad = False
        self.sampling_offsets.bias.requires_grad = False

    def freeze_attention_weights(self):
        print("Freeze attention weights")
        self.attention_weights.weight.requires_grad = False
        self.attention_weights.bias.requires_grad = False

    # ---- ZiRa -------------------------------------------------------------------------------------
    def add_zira_branches(self, value_proj: bool = True, output_proj: bool = True):
        """Attach zero-initialised re-parameterisable branches (RepZeroLinear semantics) next to
        value_proj / output_proj.  Parameter names contain "adapter", which is what the reference's
        ``before_train`` un-freezes (groundingdino_dual_zero_rep_branch.py:722-734)."""
        dev, dt = self.value_proj.weight.device, self.value_proj.weight.dtype
        if value_proj and self.value_proj_adapter is None:
            self.value_proj_adapter = RepZeroLinear(self.embed_dim, self.embed_dim).to(device=dev, dtype=dt)
        if output_proj and self.output_proj_adapter is None:
            self.output_proj_adapter = RepZeroLinear(self.embed_dim, self.embed_dim).to(device=dev, dtype=dt)
        return self

    def _project(self, x, base, adapter):
        """``base(x)`` plus, when present, the ZiRa branch: train ``base(x) + s*branch(x) + freeze(x)``
        (+ zero-inter loss), eval ``base(x) + freeze(x)``  (caller-side sum as in
        groundingdino_dual_zero_rep_branch.py:459-462)."""
        if adapter is None:
            return F.linear(x, base.weight, base.bias), None
        return adapter.forward_folded(x, base.weight, base.bias)

    # ---- fused 16-bit path (tcgen05 GEMMs with fused epilogues; see fused.py) ---------------------
    fused_enabled = True  # class-level switch used by tests / benchmarks for A/B runs

    def _use_fused(self, value, reference_points):
        if not (self.fused_enabled and value.is_cuda and fused.supported(self.embed_dim, self.num_heads, self.num_levels,
                                                                        self.num_points, value.dtype)):
            return False
        if reference_points.shape[-1] not in (2, 4) or reference_points.requires_grad:
            return False
        # the 16-bit GEMMs read raw parameter words: fp32 parameters beside 16-bit activations take the generic path
        return all(p.dtype == value.dtype for p in self.parameters())

    def _effective(self, base, adapter):
        """Eval-mode weights of a projection: pretrained + accumulated soft-frozen ZiRa weights
        (the un-merged branch is ignored in eval, groundingdino_dual_zero_rep_branch.py:126-127)."""
        if adapter is None:
            return base.weight, base.bias
        return base.weight + adapter.freeze_linear.weight, base.bias + adapter.freeze_linear.bias

    def _forward_fused_zira_train(self, query, value, key_padding_mask, reference_points, spatial_shapes,
                                  level_start_index):
        """Training mode with un-merged ZiRa branches: stage-wise Functions; each augmented projection is ONE
        tcgen05 GEMM over [W_0; W_f; W_b] whose epilogue folds the three products and reduces the zero-inter loss."""
        N, Lq, C = query.shape
        S = value.shape[1]
        M, L, P = self.num_heads, self.num_levels, self.num_points
        row_mask = None
        if key_padding_mask is not None:
            row_mask = key_padding_mask.reshape(-1).to(torch.uint8).contiguous()
        losses = []

        def project(x2d, base, ad, mask):
            if ad is None:
                return fused.Linear16Function.apply(x2d, base.weight, base.bias, mask)
            y, zl = fused.ZiRaLinear16Function.apply(x2d, mask, base.weight, base.bias, ad.freeze_linear.weight,
                                                     ad.freeze_linear.bias, ad.weight, ad.bias, ad.scaling)
            losses.append(zl)
            return y

        v = project(value.reshape(N * S, C).contiguous(), self.value_proj, self.value_proj_adapter, row_mask)
        loc, aw = fused.QueryProj16Function.apply(query.reshape(N * Lq, C).contiguous(), self.sampling_offsets.weight,
                                                  self.sampling_offsets.bias, self.attention_weights.weight,
                                                  self.attention_weights.bias, reference_points, spatial_shapes, M, L, P)
        core = MultiScaleDeformableAttnFunction.apply(v.view(N, S, M, C // M), spatial_shapes, level_start_index,
                                                      loc.view(N, Lq, M, L, P, 2), aw.view(N, Lq, M, L, P), self.im2col_step)
        out = project(core.reshape(N * Lq, C), self.output_proj, self.output_proj_adapter, None)
        self.zero_inter_loss = sum(losses) if losses else None
        return out.view(N, Lq, C)

    def _use_fused32(self, value, reference_points):
        """fp32 activations and parameters: 3 x TF32 tcgen05 projections (fused.FusedMSDeformAttnFunction32).  Un-merged
        ZiRa branches in training mode keep the stage-wise fp32 path (their loss needs the branch activations)."""
        if not (self.fused_enabled and value.is_cuda and value.dtype == torch.float32 and not torch.is_autocast_enabled()
                and fused.supported32(self.embed_dim, self.num_heads, self.num_levels, self.num_points)):
            return False
        if reference_points.shape[-1] not in (2, 4) or reference_points.requires_grad:
            return False
        if self.training and (self.value_proj_adapter is not None or self.output_proj_adapter is not None):
            return False
        return all(p.dtype == torch.float32 for p in self.parameters())

    def _prepared32(self, raw):
        key = tuple((t.data_ptr(), _version_of(t), t.dtype) for t in self.parameters())
        if getattr(self, "_prep32_key", None) != key:
            self._prep32, self._prep32_key = fused.Prepared32(*raw), key
        return self._prep32

    def _raw_weights(self):
        w_v, b_v = self._effective(self.value_proj, self.value_proj_adapter)
        w_o, b_o = self._effective(self.output_proj, self.output_proj_adapter)
        return (w_v, b_v, self.sampling_offsets.weight, self.sampling_offsets.bias, self.attention_weights.weight,
                self.attention_weights.bias, w_o, b_o)

    def _prepared(self, raw=None):
        """Kernel-ready weights (fused.Prepared), rebuilt only when a parameter changed."""
        # keyed on (storage, version, dtype): an in-place update through the parameter (optimizer step, load_state_dict,
        # __rep__) bumps the version; writing through ``.data`` does NOT -- call ``invalidate_prepared()`` after that
        key = tuple((t.data_ptr(), _version_of(t), t.dtype) for t in self.parameters())
        if getattr(self, "_prep_key", None) != key:
            self._prep, self._prep_key = fused.Prepared(*(raw if raw is not None else self._raw_weights())), key
        return self._prep

    def invalidate_prepared(self):
        """Drop the cached kernel-ready weights (needed only after mutating a parameter through ``.data``)."""
        self._prep_key = self._prep32_key = None

    def _forward_fused(self, query, value, key_padding_mask, reference_points, spatial_shapes, level_start_index):
        has_branch = self.value_proj_adapter is not None or self.output_proj_adapter is not None
        if has_branch and self.training:
            return self._forward_fused_zira_train(query, value, key_padding_mask, reference_points, spatial_shapes,
                                                  level_start_index)
        raw = self._raw_weights()
        self._prepared(raw)
        row_mask = None
        if key_padding_mask is not None:
            row_mask = key_padding_mask.reshape(-1).to(torch.uint8).contiguous()
        self.zero_inter_loss = None
        return fused.FusedMSDeformAttnFunction.apply(
            query.contiguous(), value.contiguous(), row_mask, reference_points, spatial_shapes, level_start_index,
            self._prep, self.num_heads, self.num_levels, self.num_points, self.im2col_step, *raw)

    def forward(self, query: torch.Tensor, key: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                query_pos: Optional[torch.Tensor] = None, key_padding_mask: Optional[torch.Tensor] = None,
                reference_points: Optional[torch.Tensor] = None, spatial_shapes: Optional[torch.Tensor] = None,
                level_start_index: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        """Same contract as the reference forward (ms_deform_attn.py:229-355); ``key`` is ignored."""
        if value is None:
            value = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)

        bs, num_query, _ = query.shape
        bs, num_value, _ = value.shape
        assert sum(h * w for h, w in _host_shapes(spatial_shapes)) == num_value
        if not value.is_cuda:
            raise RuntimeError("MultiScaleDeformableAttention (B200) has no CPU path; got a %s tensor" % value.device)

        M, L, P = self.num_heads, self.num_levels, self.num_points
        if self._use_fused(value, reference_points):
            output = self._forward_fused(query, value, key_padding_mask, reference_points, spatial_shapes,
                                         level_start_index)
            if not self.batch_first:
                output = output.permute(1, 0, 2)
            return output
        if self._use_fused32(value, reference_points):
            raw = self._raw_weights()
            prep = self._prepared32(raw)
            row_mask = None if key_padding_mask is None else key_padding_mask.reshape(-1).to(torch.uint8).contiguous()
            self.zero_inter_loss = None
            output = fused.FusedMSDeformAttnFunction32.apply(
                query.contiguous(), value.contiguous(), row_mask, reference_points, spatial_shapes, level_start_index, prep,
                M, L, P, self.im2col_step, *raw)
            if not self.batch_first:
                output = output.permute(1, 0, 2)
            return output
        value, loss_v = self._project(value, self.value_proj, self.value_proj_adapter)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], float(0))
        value = value.view(bs, num_value, M, -1)
        if (value.dtype == torch.float32 and (L * P) % 4 == 0 and not reference_points.requires_grad
                and reference_points.shape[-1] in (2, 4) and self.fused_enabled):
            # fp32: library GEMM for [sampling_offsets | attention_weights] (one product, shared input), then the
            # elementwise tail (locations + softmax) as two kernels instead of ~10 eager ops
            w_cat = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0)
            b_cat = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0)
            raw = F.linear(query.reshape(bs * num_query, -1), w_cat, b_cat)
            loc, aw = fused.QueryPostF32Function.apply(raw, reference_points.reshape(bs * num_query, L, -1), spatial_shapes,
                                                       M, L, P)
            output = MultiScaleDeformableAttnFunction.apply(
                value.contiguous(), spatial_shapes, level_start_index, loc.view(bs, num_query, M, L, P, 2),
                aw.view(bs, num_query, M, L, P), self.im2col_step)
            output, loss_o = self._project(output, self.output_proj, self.output_proj_adapter)
            self.zero_inter_loss = None
            if loss_v is not None or loss_o is not None:
                self.zero_inter_loss = sum(x for x in (loss_v, loss_o) if x is not None)
            if not self.batch_first:
                output = output.permute(1, 0, 2)
            return output
        acc_dtype = torch.float64 if value.dtype == torch.float64 else torch.float32
        # 16-bit activations: offsets -> locations and the softmax are evaluated in fp32
        sampling_offsets = self.sampling_offsets(query).view(bs, num_query, M, L, P, 2).to(acc_dtype)
        reference_points = reference_points.to(acc_dtype)
        attention_weights = self.attention_weights(query).view(bs, num_query, M, L * P)
        attention_weights = attention_weights.softmax(-1, dtype=acc_dtype).view(bs, num_query, M, L, P)

        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            sampling_locations = (reference_points[:, :, None, :, None, :]
                                  + sampling_offsets / offset_normalizer[None, None, None, :, None, :])
        elif reference_points.shape[-1] == 4:
            sampling_locations = (reference_points[:, :, None, :, None, :2]
                                  + sampling_offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5)
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))

        output = MultiScaleDeformableAttnFunction.apply(
            value.contiguous(), spatial_shapes, level_start_index,
            sampling_locations.contiguous(), attention_weights.contiguous(), self.im2col_step)

        output, loss_o = self._project(output, self.output_proj, self.output_proj_adapter)
        self.zero_inter_loss = None
        if loss_v is not None or loss_o is not None:
            self.zero_inter_loss = sum(x for x in (loss_v, loss_o) if x is not None)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return output
