"""Drop-in for the reference's pybind module ``groundingdino._C`` (csrc/vision.cpp:53-56).

Same two functions, same argument order and meaning, same error behaviour (``RuntimeError`` for
non-contiguous / non-CUDA inputs and for ``batch % min(batch, im2col_step) != 0``,
csrc/MsDeformAttn/ms_deform_attn_cuda.cu:29-53), but backed by the sm_100a kernels behind the C ABI
in include/msda_b200.h.  Inputs are borrowed; outputs are freshly allocated tensors; kernels are
enqueued on the current CUDA stream of the inputs' device with no synchronisation.

Supported storage dtypes: float32, float64 (as the reference), plus bfloat16 and float16 (new: value
and output in 16 bit, sampling locations / attention weights and all gradients in fp32).
"""
import torch

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.bfloat16: "bf16", torch.float16: "f16"}


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, extra=()):
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)] + list(extra)
    for name, t in named:
        _require(t.is_contiguous(), "%s tensor has to be contiguous" % name)
    for name, t in named:
        _require(t.is_cuda, "%s must be a CUDA tensor" % name)
        _require(t.device == value.device, "%s must be on the same device as value" % name)
    _require(value.dtype in _SUFFIX, "ms_deform_attn not implemented for '%s'" % value.dtype)
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes and level_start_index must be int64")
    aux = torch.float64 if value.dtype == torch.float64 else torch.float32
    _require(sampling_loc.dtype == aux and attn_weight.dtype == aux,
             "sampling_loc and attn_weight must be %s for %s value" % (aux, value.dtype))
    _require(value.dim() == 4 and sampling_loc.dim() == 6 and attn_weight.dim() == 5, "bad tensor rank")
    N, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    _require(tuple(sampling_loc.shape) == (N, Lq, M, L, P, 2), "sampling_loc shape mismatch")
    _require(tuple(attn_weight.shape) == (N, Lq, M, L, P), "attn_weight shape mismatch")
    _require(level_start_index.numel() == L, "level_start_index length mismatch")
    step = min(N, int(im2col_step))
    _require(N == 0 or (step > 0 and N % step == 0), "batch(%d) must divide im2col_step(%d)" % (N, step))
    return N, S, M, D, L, Lq, P


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> Tensor [N, Lq, M*D] (csrc/MsDeformAttn/ms_deform_attn_cuda.cu:21-81)."""
    N, S, M, D, L, Lq, P = _common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                          im2col_step)
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    if out.numel() == 0:
        return out
    fn = getattr(_lib.lib(), "msda_forward_" + _SUFFIX[value.dtype])
    with torch.cuda.device(value.device):
        rc = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(), _stream(value.device))
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def backward_workspace(value, N, M, D, Lq, P):
    """Scratch tensor for msda_backward_16_ws, or None when the range-planned scatter does not apply (fp32 / fp64 storage,
    other head sizes, tuning key ``bwd_mma_levels`` == 0)."""
    if value.dtype not in (torch.bfloat16, torch.float16) or D != 32 or P != 4 or _lib.get_tuning("bwd_mma_levels") <= 0:
        return None
    n = _lib.lib().msda_backward_workspace_bytes(N, M, Lq) // 8
    return torch.empty(max(n, 1), dtype=torch.int64, device=value.device)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight] (ms_deform_attn_cuda.cu:84-154).

    For 16-bit ``value`` the three gradients are fp32 (accumulation type); the autograd Function
    casts grad_value back to the storage dtype."""
    N, S, M, D, L, Lq, P = _common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                          im2col_step, extra=[("grad_output", grad_output)])
    _require(grad_output.dtype == value.dtype and grad_output.numel() == N * Lq * M * D, "grad_output mismatch")
    gdt = torch.float64 if value.dtype == torch.float64 else torch.float32
    grad_value = torch.empty(value.shape, dtype=gdt, device=value.device)
    grad_loc = torch.empty(sampling_loc.shape, dtype=gdt, device=value.device)
    grad_aw = torch.empty(attn_weight.shape, dtype=gdt, device=value.device)
    if N == 0 or Lq == 0:
        return [grad_value.zero_(), grad_loc, grad_aw]
    ws = backward_workspace(value, N, M, D, Lq, P)
    if ws is not None:   # 16-bit storage with the range-planned tensor-memory scatter enabled: scratch for its hit masks
        with torch.cuda.device(value.device):
            rc = _lib.lib().msda_backward_16_ws(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(), 0, 0,
                                                N, S, M, D, L, Lq, P, grad_value.data_ptr(), grad_loc.data_ptr(),
                                                grad_aw.data_ptr(), 0, 1, 1 if value.dtype == torch.float16 else 0,
                                                ws.data_ptr(), ws.numel() * 8, _stream(value.device))
        _lib.check(rc, "ms_deform_attn_backward")
        return [grad_value, grad_loc, grad_aw]
    fn = getattr(_lib.lib(), "msda_backward_" + _SUFFIX[value.dtype])
    with torch.cuda.device(value.device):
        rc = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), grad_output.data_ptr(), N, S, M, D, L, Lq, P, grad_value.data_ptr(),
                grad_loc.data_ptr(), grad_aw.data_ptr(), 1, _stream(value.device))
    _lib.check(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc, grad_aw]
