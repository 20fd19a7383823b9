"""Fused pieces of the encoder layer around MultiScaleDeformableAttention."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class AddLayerNormFunction(Function):
    """``LayerNorm(x + r)`` over the last dimension, 16-bit activations, one kernel per direction
    (reference: ``src = self.norm1(src + self.dropout1(src2))``, transformer_for_adapter.py:901-902)."""

    @staticmethod
    def forward(ctx, x, r, weight, bias, eps):
        shape = x.shape
        C = shape[-1]
        x2, r2 = x.reshape(-1, C).contiguous(), r.reshape(-1, C).contiguous()
        R = x2.shape[0]
        z, y = torch.empty_like(x2), torch.empty_like(x2)
        mean = torch.empty(R, dtype=torch.float32, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        g32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(x.device):
            rc = _lib.lib().msda_add_layernorm_fwd_16(x2.data_ptr(), r2.data_ptr(), g32.data_ptr(), b32.data_ptr(), R, C, float(eps),
                                                      z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                      1 if x.dtype == torch.float16 else 0, _stream(x))
        _lib.check(rc, "msda_add_layernorm_fwd_16")
        ctx.save_for_backward(z, g32, mean, rstd)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, g32, mean, rstd = ctx.saved_tensors
        R, C = z.shape
        dy2 = dy.reshape(R, C).contiguous()
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            rc = _lib.lib().msda_add_layernorm_bwd_16(dy2.data_ptr(), z.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                      R, C, dz.data_ptr(), 1 if z.dtype == torch.float16 else 0, _stream(z))
        _lib.check(rc, "msda_add_layernorm_bwd_16")
        dw = db = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:   # affine parameters are frozen in the ZiRa configuration
            xhat = (z.float() - mean[:, None]) * rstd[:, None]
            dw = (dy2.float() * xhat).sum(0).to(dy.dtype) if ctx.needs_input_grad[2] else None
            db = dy2.float().sum(0).to(dy.dtype) if ctx.needs_input_grad[3] else None
        dz = dz.view(ctx.shape)
        return dz, dz, dw, db, None


def add_layer_norm(x, r, norm, dropout_p=0.0, training=False):
    """``norm(x + dropout(r))`` -- fused when possible (CUDA, 16-bit, no active dropout, C % 8 == 0, C <= 1024)."""
    C = x.shape[-1]
    if (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and r.dtype == x.dtype and not (training and dropout_p > 0)
            and C % 8 == 0 and C <= 1024 and norm.elementwise_affine and norm.bias is not None):
        return AddLayerNormFunction.apply(x, r, norm.weight, norm.bias, norm.eps)
    if training and dropout_p > 0:
        r = torch.nn.functional.dropout(r, dropout_p, True)
    return norm(x + r)
