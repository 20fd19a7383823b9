"""B200-native multi-scale deformable attention for ZiRa-GroundingDINO (drop-in for the reference's
``groundingdino/models/GroundingDINO/ms_deform_attn.py`` + ``groundingdino._C``)."""
from . import _C  # noqa: F401
from .ms_deform_attn import (MultiScaleDeformableAttention, MultiScaleDeformableAttnFunction,  # noqa: F401
                             multi_scale_deformable_attn_pytorch)
from .zira import RepZeroConv2d, RepZeroLinear, merge_all  # noqa: F401
from .input_proj import ZiRaInputProj  # noqa: F401
from .encoder import DeformableEncoder, DeformableTransformerEncoderLayer  # noqa: F401
from .decoder import DeformableTransformerDecoderLayer  # noqa: F401
from .fuse_modules import BiAttentionBlock, BiMultiHeadAttention  # noqa: F401

__all__ = ["MultiScaleDeformableAttention", "MultiScaleDeformableAttnFunction", "RepZeroLinear", "RepZeroConv2d",
           "ZiRaInputProj", "BiAttentionBlock", "BiMultiHeadAttention", "DeformableTransformerEncoderLayer", "DeformableTransformerDecoderLayer", "DeformableEncoder",
           "merge_all", "_C"]
