"""GPU: row N3 of SURVEY.md 8(f) -- ZiRa conv adapters beside input_proj + GroupNorm on channels-last rows -- against the
oracle restatement (oracle.input_proj_level, pinned to the reference fixtures zira_rep_conv*.npz) evaluated in fp64 on
the same 16-bit-rounded operands.  Tolerances: one 16-bit rounding on 16-bit outputs (bf16 2^-8, f16 2^-11) plus the
rounding of the 16-bit intermediate the GroupNorm reads; 2e-2 relative (Frobenius) on gradients."""
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, dtype, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to(dtype).to(DEV)


@pytest.mark.parametrize("dtype,eps", [(torch.bfloat16, 2 ** -8), (torch.float16, 2 ** -11)])
@pytest.mark.parametrize("N,HW,C,G", [(2, 1000, 256, 32), (1, 16700, 256, 32), (3, 37, 64, 4), (2, 273, 16, 4), (1, 5, 512, 32)])
def test_group_norm_rows_vs_fp64(dtype, eps, N, HW, C, G):
    from ziragroundingdino_b200.layer_ops import GroupNormRowsFunction
    x = _rand((N, HW, C), dtype, 1, 1.5, 0.7).requires_grad_(True)
    w, b = _rand((C,), dtype, 2, 0.3, 1.0).requires_grad_(True), _rand((C,), dtype, 3, 0.3).requires_grad_(True)
    S = HW + 11
    gfull = _rand((N, S, C), dtype, 4)                   # the gradient arrives as a level slice of [N, S, C]
    y = GroupNormRowsFunction.apply(x, w, b, G, 1e-5)
    y.backward(gfull[:, 11:])
    xd, wd, bd = (t.detach().double().cpu().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.group_norm(xd.transpose(1, 2), G, wd, bd, 1e-5).transpose(1, 2)
    ref.backward(gfull[:, 11:].double().cpu())
    assert y.dtype == dtype and (y.double().cpu() - ref.detach()).abs().max().item() <= eps * ref.abs().max().item() * 1.05 + 1e-6
    assert rel_err(x.grad.double().cpu(), xd.grad) < 2e-2
    assert rel_err(w.grad.double().cpu(), wd.grad) < 2e-2 and rel_err(b.grad.double().cpu(), bd.grad) < 2e-2


def test_group_norm_rows_rejects_bad_shapes():
    from ziragroundingdino_b200 import _lib
    L = _lib.lib()
    x = torch.zeros(1, 4, 24, dtype=torch.bfloat16, device=DEV)
    f = torch.zeros(24, device=DEV)
    mr = torch.zeros(1, 3, 2, device=DEV)
    sc = torch.zeros(1, 24, 2, dtype=torch.float64, device=DEV)
    # C/8 = 3 does not divide 256 -> unsupported, not a silent wrong answer
    rc = L.msda_group_norm_fwd_16(x.data_ptr(), 96, f.data_ptr(), f.data_ptr(), 1, 4, 24, 3, 1e-5, x.data_ptr(), 96,
                                  mr.data_ptr(), sc.data_ptr(), 0, 0)
    assert rc != 0
    rc = L.msda_group_norm_fwd_16(0, 96, f.data_ptr(), f.data_ptr(), 1, 4, 24, 3, 1e-5, x.data_ptr(), 96, mr.data_ptr(),
                                  sc.data_ptr(), 0, 0)
    assert rc != 0


def _oracle_level(O, x_rows, hw, conv, norm, ad, training):
    """oracle.input_proj_level in fp64 on CPU from the module's (16-bit) parameters; returns rows [N, H'W', C]."""
    N, HW, Cin = x_rows.shape
    d = lambda t: t.detach().double().cpu()
    x = d(x_rows).reshape(N, hw[0], hw[1], Cin).permute(0, 3, 1, 2)
    leaves = [d(p).requires_grad_(True) for p in (ad.weight, ad.bias, ad.scaling, ad.freeze_conv.weight, ad.freeze_conv.bias)]
    src, loss = O.input_proj_level(x, d(conv.weight), d(conv.bias), d(norm.weight), d(norm.bias), norm.num_groups, leaves,
                                   training, stride=conv.stride, padding=conv.padding, eps=norm.eps)
    return src.flatten(2).transpose(1, 2), loss, leaves


@pytest.mark.parametrize("cin,k,hw", [(192, 1, (40, 67)), (384, 1, (20, 34)), (768, 1, (10, 17)), (768, 3, (10, 17)), (256, 3, (5, 9))])
def test_input_proj_level_fused_vs_oracle(cin, k, hw):
    """One level, training mode, bf16: fused ZiRa GEMM (K = C_in or 9*C_in after im2col) + row GroupNorm vs the oracle;
    output, zero-inter loss and the gradients of all five branch tensors."""
    from oracle import msda_oracle as O
    import ziragroundingdino_b200 as zb
    torch.manual_seed(7)
    dt = torch.bfloat16
    stride, pad = (1, 0) if k == 1 else (2, 1)
    conv = torch.nn.Conv2d(cin, 256, k, stride, pad)
    norm = torch.nn.GroupNorm(32, 256)
    ad = zb.RepZeroConv2d(cin, 256, kernel_size=k, stride=stride, padding=pad)
    with torch.no_grad():
        sc = (cin * k * k) ** -0.5
        conv.weight.normal_(0, sc); conv.bias.normal_(0, 0.1)
        norm.weight.normal_(1, 0.2); norm.bias.normal_(0, 0.2)
        ad.weight.normal_(0, 0.5 * sc); ad.bias.normal_(0, 0.1)
        ad.freeze_conv.weight.normal_(0, 0.5 * sc); ad.freeze_conv.bias.normal_(0, 0.1)
        ad.scaling.fill_(0.3)
    conv, norm, ad = conv.to(DEV, dt), norm.to(DEV, dt), ad.to(DEV, dt)
    ad.train()
    N = 2
    x_rows = _rand((N, hw[0] * hw[1], cin), dt, 8)
    from ziragroundingdino_b200.layer_ops import group_norm_rows
    y, out_hw, loss = ad.forward_folded_rows(x_rows, hw, conv)
    src = group_norm_rows(y, norm)
    g = _rand(tuple(src.shape), dt, 9)
    (src.float() * g.float()).sum().add(loss.float() * 50.0).backward()
    ref, ref_loss, leaves = _oracle_level(O, x_rows, hw, conv, norm, ad, True)
    ((ref * g.double().cpu()).sum() + ref_loss * 50.0).backward()
    assert tuple(src.shape) == tuple(ref.shape)
    # src is O(1) after the normalisation; its input y was rounded to bf16 once (relative 2^-9 of |y|, amplified by rstd*gamma)
    assert (src.double().cpu() - ref.detach()).abs().max().item() < 4 * 2 ** -8 * ref.abs().max().item()
    assert abs(float(loss) - float(ref_loss)) <= 1e-2 * float(ref_loss)
    names = ["weight", "bias", "scaling", "freeze_conv.weight", "freeze_conv.bias"]
    got = [ad.weight, ad.bias, ad.scaling, ad.freeze_conv.weight, ad.freeze_conv.bias]
    for n, a, b in zip(names, got, leaves):
        assert a.grad is not None and a.grad.shape == a.shape, n
        assert rel_err(a.grad.double().cpu(), b.grad) < (6e-2 if n == "scaling" else 2e-2), n


@pytest.mark.parametrize("training", [True, False])
def test_zira_input_proj_module_vs_eager(training):
    """Whole ZiRaInputProj (Swin-T channel widths, 4 levels) in bf16: fused rows path vs its own eager NCHW path
    (library conv2d + GroupNorm, the reference's op sequence), then merge equivalence (after_train, :739-745)."""
    import ziragroundingdino_b200 as zb
    torch.manual_seed(11)
    m = zb.ZiRaInputProj().to(DEV, torch.bfloat16)
    with torch.no_grad():
        for a in m.input_proj_conv_adapter:
            sc = (a.weight[0].numel()) ** -0.5
            a.weight.normal_(0, 0.3 * sc); a.freeze_conv.weight.normal_(0, 0.3 * sc); a.freeze_conv.bias.normal_(0, 0.05)
    m.train(training)
    hw = [(25, 42), (13, 21), (7, 11)]
    feats = [_rand((2, c, h, w), torch.bfloat16, 20 + i) for i, (c, (h, w)) in enumerate(zip((192, 384, 768), hw))]
    outs, loss = m(feats)
    want, wl = [], 0.0
    for l in range(4):
        x = feats[l] if l < 3 else feats[-1]
        a, zl = m.input_proj_conv_adapter[l](x)
        want.append(m.input_proj[l][1](m.input_proj[l][0](x) + a))
        wl = wl + zl.float()
    assert [tuple(o.shape) for o in outs] == [tuple(w.shape) for w in want] and outs[3].shape[-2:] == (4, 6)
    for o, w in zip(outs, want):
        assert (o.float() - w.float()).abs().max().item() < 6 * 2 ** -8 * w.float().abs().max().item()
    assert abs(float(loss) - float(wl)) <= 2e-2 * float(wl) + 1e-12
    if training:
        m.eval()
        zb.merge_all(m)
        merged, _ = m(feats)
        for o, w in zip(merged, want):
            assert (o.float() - w.float()).abs().max().item() < 6 * 2 ** -8 * w.float().abs().max().item()


def test_input_proj_fp32_master_adapters_with_bf16_activations():
    """Mixed precision the way a trainer would run it: frozen input_proj in bf16, trainable adapter parameters kept in
    fp32 (cast to bf16 on the way into the fused GEMM, gradients arrive back in fp32)."""
    import ziragroundingdino_b200 as zb
    torch.manual_seed(13)
    m = zb.ZiRaInputProj().to(DEV)
    m.input_proj.to(torch.bfloat16)
    with torch.no_grad():
        for a in m.input_proj_conv_adapter:
            sc = (a.weight[0].numel()) ** -0.5
            a.weight.normal_(0, 0.3 * sc); a.freeze_conv.weight.normal_(0, 0.3 * sc)
    for n, p in m.named_parameters():
        p.requires_grad_("adapter" in n)
    m.train()
    hw = [(25, 42), (13, 21), (7, 11)]
    rows = [_rand((2, h * w, c), torch.bfloat16, 30 + i) for i, (c, (h, w)) in enumerate(zip((192, 384, 768), hw))]
    src, shapes, loss = m.forward_rows(rows, hw)
    assert src.dtype == torch.bfloat16 and shapes == [(25, 42), (13, 21), (7, 11), (4, 6)]
    (src.float().square().mean() + 0.1 * loss.float()).backward()
    for n, p in m.named_parameters():
        if "adapter" in n:
            assert p.grad is not None and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all(), n
    # same module with bf16 adapters: the fp32-master run must agree to bf16 rounding of the parameters
    m2 = zb.ZiRaInputProj().to(DEV)
    m2.load_state_dict(m.state_dict())
    m2 = m2.to(torch.bfloat16).train()
    src2, _, loss2 = m2.forward_rows(rows, hw)
    assert (src.float() - src2.float()).abs().max().item() < 0.1 and abs(float(loss) - float(loss2)) < 2e-2 * float(loss2)


def test_five_level_training_backward_through_im2col_level():
    """with num_feature_levels = backbone levels + 2 the last level reads the previous PROJECTED level, which
    depends on trainable adapters, so the im2col'd 3x3 level (K = 9*256 = 2304 > the tcgen05 GEMM's 2048-column limit)
    needs d(input) in backward: it must run (library product for that one dgrad), not raise."""
    import ziragroundingdino_b200 as zb
    torch.manual_seed(17)
    m = zb.ZiRaInputProj((192, 384, 768), 256, 5).to(DEV, torch.bfloat16)
    with torch.no_grad():
        for a in m.input_proj_conv_adapter:
            sc = (a.weight[0].numel()) ** -0.5
            a.weight.normal_(0, 0.3 * sc); a.freeze_conv.weight.normal_(0, 0.3 * sc)
    for n, p in m.named_parameters():
        p.requires_grad_("adapter" in n)
    m.train()
    hw = [(25, 42), (13, 21), (7, 11)]
    rows = [_rand((2, h * w, c), torch.bfloat16, 40 + i) for i, (c, (h, w)) in enumerate(zip((192, 384, 768), hw))]
    src, shapes, loss = m.forward_rows(rows, hw)
    assert shapes == [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
    # a loss on the LAST level only: its gradient reaches level 3's adapter solely through the K = 2304 dgrad.  A fixed random
    # linear functional (mean-square of a GroupNorm output is ~constant: its gradient would be round-off)
    gy = _rand((2, 6, 256), torch.bfloat16, 49)
    last = src[:, -6:]
    (last.float() * gy.float()).sum().backward()
    g3 = m.input_proj_conv_adapter[3].weight.grad
    assert g3 is not None and torch.isfinite(g3).all() and g3.abs().max() > 0
    # eager NCHW reference of the same two levels in fp64 for the gradient of level 3's branch weight
    d = lambda t: t.detach().double()
    a3, a4 = m.input_proj_conv_adapter[3], m.input_proj_conv_adapter[4]
    w3 = d(a3.weight).requires_grad_(True)
    x = d(rows[2]).transpose(1, 2).reshape(2, 768, 7, 11)
    F_ = torch.nn.functional
    y3 = F_.conv2d(x, d(m.input_proj[3][0].weight), d(m.input_proj[3][0].bias), stride=2, padding=1) \
        + d(a3.scaling) * F_.conv2d(x, w3, d(a3.bias), stride=2, padding=1) + F_.conv2d(x, d(a3.freeze_conv.weight), d(a3.freeze_conv.bias), stride=2, padding=1)
    s3 = F_.group_norm(y3, 32, d(m.input_proj[3][1].weight), d(m.input_proj[3][1].bias), 1e-5)
    s3 = s3.detach().to(torch.bfloat16).double() + (s3 - s3.detach())   # value: the bf16-rounded stored level; gradient: identity
    y4 = F_.conv2d(s3, d(m.input_proj[4][0].weight), d(m.input_proj[4][0].bias), stride=2, padding=1) \
        + d(a4.scaling) * F_.conv2d(s3, d(a4.weight), d(a4.bias), stride=2, padding=1) + F_.conv2d(s3, d(a4.freeze_conv.weight), d(a4.freeze_conv.bias), stride=2, padding=1)
    s4 = F_.group_norm(y4, 32, d(m.input_proj[4][1].weight), d(m.input_proj[4][1].bias), 1e-5)
    (s4.flatten(2).transpose(1, 2) * gy.double()).sum().backward()
    assert rel_err(g3.double().cpu(), w3.grad.cpu()) < 5e-2
