"""ZiRa re-parameterisable zero-initialised branches.

Semantics follow the reference's ``RepZeroLinear`` (groundingdino_dual_zero_rep_branch.py:105-135;
constants :62-64; merge trigger ``after_train`` :739-745):

  construct  branch weight := 1e-8, branch bias keeps nn.Linear's default init, scaling := 0.1,
             freeze_linear weight/bias := 0
  train      branch = scaling * (W_b x + b_b);  out = branch + (W_f x + b_f)
             returns (out, SmoothL1(branch, 0) + SmoothL1(out, 0))
  eval       returns (W_f x + b_f, 0)           -- an un-merged branch is ignored in eval
  __rep__    W_f += scaling*W_b ; b_f += scaling*b_b ; scaling := 0.1 ; W_b := 1e-8 ; b_b := 1e-8

Parameter names (``weight``, ``bias``, ``scaling``, ``freeze_linear.weight``, ``freeze_linear.bias``)
match the reference so its optimiser rule (lr factor for names containing "freeze",
test_odinw13_softfreeze/for_train/*.py:24) applies unchanged.

``forward_folded`` is the extension BASELINE.json asks for: the branch sits beside a frozen
``nn.Linear`` (value_proj / output_proj of MultiScaleDeformableAttention) and the three weight sets
(pretrained W_0, soft-frozen W_f, fresh s*W_b) are contracted in ONE pass over the activation.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

zero_value = 1e-8
lan_scale = 0.1
vis_scale = 0.1


def _folded_rows(x_rows, training, w0, b0, wf, bf, wb, bb, scaling):
    """Rows [.., K] -> (base + adapter, loss) for weight matrices [F, K]: the shared body of the linear and the
    im2col'd convolution fold.  Training on CUDA 16-bit = ONE tcgen05 GEMM over [W_0; W_f; W_b] (fused.py)."""
    Fo, K = w0.shape
    if not training:
        return F.linear(x_rows, w0 + wf, b0 + bf), None
    if (x_rows.is_cuda and x_rows.dtype in (torch.bfloat16, torch.float16) and K % 64 == 0 and Fo % 64 == 0 and 3 * Fo <= 2048):
        from . import fused
        x2d = x_rows.reshape(-1, K).contiguous()
        cast = lambda t: t if t.dtype == x2d.dtype else t.to(x2d.dtype)
        y, loss = fused.ZiRaLinear16Function.apply(x2d, None, *(cast(t) for t in (w0, b0, wf, bf, wb, bb, scaling)))
        return y.view(*x_rows.shape[:-1], Fo), loss
    branch = scaling * F.linear(x_rows, wb, bb)
    adapter_out = branch + F.linear(x_rows, wf, bf)
    loss = (F.smooth_l1_loss(branch, torch.zeros_like(branch)) + F.smooth_l1_loss(adapter_out, torch.zeros_like(adapter_out)))
    return F.linear(x_rows, w0, b0) + adapter_out, loss


class RepZeroConv2d(nn.Conv2d):
    """The reference's convolutional ZiRa branch (groundingdino_dual_zero_rep_branch.py:64-103): same constructor,
    parameter names (``weight``, ``bias``, ``scaling``, ``freeze_conv.weight``, ``freeze_conv.bias``), train / eval
    outputs and ``__rep__`` merge as there.  It sits beside ``input_proj[l][0]`` (:292-302, :492-493)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode="zeros", device=None, dtype=None, zero_value=zero_value):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode, device, dtype)
        self.scaling = nn.parameter.Parameter(torch.ones(1, device=device, dtype=dtype) * vis_scale)
        nn.init.constant_(self.weight, val=zero_value)
        if self.bias is not None:
            nn.init.constant_(self.bias, val=zero_value)
        self.freeze_conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                                     padding_mode, device, dtype)
        nn.init.constant_(self.freeze_conv.weight, val=0.0)
        if self.bias is not None:
            nn.init.constant_(self.freeze_conv.bias, val=0.0)
        self.zero_inter_loss = torch.nn.SmoothL1Loss(reduction="mean")

    def forward(self, input):
        """Stand-alone use, exactly the reference contract: NCHW in, ``(output, zero_inter_loss)`` out."""
        if self.training:
            branch_output = self.scaling * super().forward(input)
            output = branch_output + self.freeze_conv(input)
            loss = (self.zero_inter_loss(branch_output, torch.zeros_like(branch_output))
                    + self.zero_inter_loss(output, torch.zeros_like(output)))
            return output, loss
        return self.freeze_conv(input), torch.zeros(1).to(input)

    def rows_foldable(self, base_conv):
        """True when conv == GEMM over channels-last rows (possibly after im2col): what forward_folded_rows needs."""
        return (self.groups == 1 and self.dilation == (1, 1) and self.padding_mode == "zeros" and self.bias is not None
                and isinstance(base_conv, nn.Conv2d) and base_conv.bias is not None
                and base_conv.weight.shape == self.weight.shape and base_conv.stride == self.stride
                and base_conv.padding == self.padding and base_conv.dilation == self.dilation and base_conv.groups == 1)

    def forward_folded_rows(self, x_rows, hw, base_conv):
        """``base_conv(x) + self(x)[0]`` and the zero-inter loss on channels-last rows.

        x_rows [N, H*W, C_in] (hw = (H, W)); returns (rows [N, H_out*W_out, C_out], (H_out, W_out), loss or None).
        A 1x1 convolution is a GEMM over the rows as they are; a k x k one is im2col'd first (the only such level of the
        reference is the 13x21 extra level, :269-276).  Both then run as one fused GEMM in training (see _folded_rows)."""
        N, HW, Cin = x_rows.shape
        H, W = hw
        kh, kw = self.kernel_size
        Fo = self.out_channels
        as_rows = lambda w: w.reshape(Fo, -1)
        if (kh, kw) == (1, 1) and self.stride == (1, 1) and self.padding == (0, 0):
            rows, out_hw = x_rows, (H, W)
        else:
            # im2col straight from the channels-last map: pad, then a strided window view [N, Ho, Wo, kh, kw, C] copied
            # once; K runs (ky, kx, c), so the weights are permuted to match instead of transposing the activation
            ph, pw = self.padding
            sh_, sw_ = self.stride
            Ho, Wo = (H + 2 * ph - kh) // sh_ + 1, (W + 2 * pw - kw) // sw_ + 1
            xp = F.pad(x_rows.reshape(N, H, W, Cin), (0, 0, pw, pw, ph, ph))
            Hp, Wp = H + 2 * ph, W + 2 * pw
            win = xp.as_strided((N, Ho, Wo, kh, kw, Cin), (Hp * Wp * Cin, sh_ * Wp * Cin, sw_ * Cin, Wp * Cin, Cin, 1))
            rows, out_hw = win.reshape(N, Ho * Wo, kh * kw * Cin), (Ho, Wo)
            as_rows = lambda w: w.permute(0, 2, 3, 1).reshape(Fo, -1)
        cast = (lambda t: t) if self.weight.dtype == x_rows.dtype else (lambda t: t.to(x_rows.dtype))
        y, loss = _folded_rows(rows, self.training, as_rows(base_conv.weight), base_conv.bias,
                               cast(as_rows(self.freeze_conv.weight)), cast(self.freeze_conv.bias), cast(as_rows(self.weight)),
                               cast(self.bias), cast(self.scaling))
        return y, out_hw, loss

    def __rep__(self):
        with torch.no_grad():
            self.freeze_conv.weight.data = self.weight.data * self.scaling + self.freeze_conv.weight.data
            if self.bias is not None:
                self.freeze_conv.bias.data = self.bias.data * self.scaling + self.freeze_conv.bias.data
        self.scaling = nn.parameter.Parameter(torch.ones(1).to(self.weight.data) * vis_scale)
        nn.init.constant_(self.weight, val=zero_value)
        if self.bias is not None:
            nn.init.constant_(self.bias, val=zero_value)


class RepZeroLinear(nn.Linear):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None) -> None:
        super().__init__(in_features, out_features, bias, device, dtype)
        self.scaling = nn.parameter.Parameter(torch.ones(1, device=device, dtype=dtype) * lan_scale)
        nn.init.constant_(self.weight, val=zero_value)
        self.freeze_linear = nn.Linear(in_features, out_features, bias, device, dtype)
        nn.init.constant_(self.freeze_linear.weight, val=0.0)
        if self.bias is not None:
            nn.init.constant_(self.freeze_linear.bias, val=0.0)
        self.zero_inter_loss = torch.nn.SmoothL1Loss(reduction="mean")

    def forward(self, input):
        """Stand-alone use, exactly the reference contract: returns ``(output, zero_inter_loss)``."""
        if self.training:
            branch_output = self.scaling * F.linear(input, self.weight, self.bias)
            output = branch_output + self.freeze_linear(input)
            loss = (self.zero_inter_loss(branch_output, torch.zeros_like(branch_output))
                    + self.zero_inter_loss(output, torch.zeros_like(output)))
            return output, loss
        return self.freeze_linear(input), torch.zeros(1).to(input)

    def forward_folded(self, input, base_weight, base_bias):
        """``F.linear(input, base_weight, base_bias) + self(input)[0]`` and the zero-inter loss.

        Weight-space fold: one [out, 2*in]... GEMM over the activation is not needed -- the three weight
        sets are summed (a 256x256 axpy) and applied once; the loss terms need the branch and the
        adapter output separately, so in training the branch and freeze products are formed too.
        """
        if not self.training:
            w = base_weight + self.freeze_linear.weight
            b = None if base_bias is None else base_bias + self.freeze_linear.bias
            return F.linear(input, w, b), None
        if (input.is_cuda and input.dtype in (torch.bfloat16, torch.float16) and self.bias is not None and base_bias is not None
                and self.in_features % 64 == 0 and self.out_features % 64 == 0 and 3 * self.out_features <= 2048):
            # one tcgen05 GEMM over [W_0; W_f; W_b] with the fold and the loss reduction in its epilogue (fused.py)
            from . import fused
            x2d = input.reshape(-1, self.in_features).contiguous()
            cast = (lambda t: t) if self.weight.dtype == x2d.dtype else (lambda t: t.to(x2d.dtype))   # fp32 master adapters
            y, loss = fused.ZiRaLinear16Function.apply(x2d, None, cast(base_weight), cast(base_bias), cast(self.freeze_linear.weight),
                                                       cast(self.freeze_linear.bias), cast(self.weight), cast(self.bias),
                                                       cast(self.scaling))
            return y.view(*input.shape[:-1], self.out_features), loss
        branch = self.scaling * F.linear(input, self.weight, self.bias)
        adapter_out = branch + self.freeze_linear(input)
        loss = (self.zero_inter_loss(branch, torch.zeros_like(branch))
                + self.zero_inter_loss(adapter_out, torch.zeros_like(adapter_out)))
        return F.linear(input, base_weight, base_bias) + adapter_out, loss

    def __rep__(self):
        with torch.no_grad():
            self.freeze_linear.weight.data = self.weight.data * self.scaling + self.freeze_linear.weight.data
            if self.bias is not None:
                self.freeze_linear.bias.data = self.bias.data * self.scaling + self.freeze_linear.bias.data
        self.scaling = nn.parameter.Parameter(torch.ones(1).to(self.weight.data) * lan_scale)
        nn.init.constant_(self.weight, val=zero_value)
        if self.bias is not None:
            nn.init.constant_(self.bias, val=zero_value)


def merge_all(module: nn.Module):
    """The reference's ``after_train`` (groundingdino_dual_zero_rep_branch.py:739-745): merge every branch."""
    for m in module.modules():
        if hasattr(m, "__rep__"):
            m.__rep__()
