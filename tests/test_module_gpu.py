"""GPU: the drop-in module against the reference module's own outputs (golden fixtures produced by
running the reference's MultiScaleDeformableAttention on CPU, tests/golden/make_golden.py) and
against the oracle's restatement of the module; ZiRa branch semantics (train / eval / merge)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu


def _load_module(g, dtype, dev):
    import ziragroundingdino_b200 as zb
    C, M, L, P, bf = (int(x) for x in g["cfg"])
    m = zb.MultiScaleDeformableAttention(C, M, L, P, batch_first=bool(bf)).double()  # fixtures are fp64
    m.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    return m.to(device=dev, dtype=dtype), (C, M, L, P, bool(bf))


def _inputs(g, dtype, dev, bf):
    q = torch.from_numpy(g["query"]).to(device=dev, dtype=dtype)
    v = torch.from_numpy(g["value"]).to(device=dev, dtype=dtype)
    if not bf:
        q, v = q.transpose(0, 1).contiguous(), v.transpose(0, 1).contiguous()
    q.requires_grad_(True); v.requires_grad_(True)
    refp = torch.from_numpy(g["reference_points"]).to(device=dev, dtype=dtype)
    sh = torch.from_numpy(g["shapes"]).to(dev)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    mask = torch.from_numpy(g["mask"]).to(dev) if "mask" in g else None
    return q, v, refp, sh, lsi, mask


@pytest.mark.parametrize("name", ["module_enc", "module_dec", "module_seqfirst", "module_d32"])
@pytest.mark.parametrize("dtype,tol_out,tol_grad", [(torch.float64, 1e-11, 1e-10), (torch.float32, 2e-5, 1e-4)])
def test_module_matches_reference_fixture(name, dtype, tol_out, tol_grad):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    m, (C, M, L, P, bf) = _load_module(g, dtype, dev)
    q, v, refp, sh, lsi, mask = _inputs(g, dtype, dev, bf)
    out = m(query=q, value=v, key_padding_mask=mask, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
    assert out.shape == torch.Size(g["out"].shape)
    assert np.abs(out.detach().cpu().double().numpy() - g["out"]).max() < tol_out
    out.backward(torch.from_numpy(g["grad_out"]).to(device=dev, dtype=dtype))
    gq = q.grad if bf else q.grad.transpose(0, 1)
    gvi = v.grad if bf else v.grad.transpose(0, 1)
    assert rel_err(gq.cpu(), g["grad_query"]) < tol_grad
    assert rel_err(gvi.cpu(), g["grad_value_in"]) < tol_grad
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), g["pgrad." + k]) < tol_grad, k


def test_module_bf16_within_1e2():
    """bf16 module vs fp64 reference fixture evaluated on the same (bf16-rounded) weights and inputs."""
    dev = torch.device("cuda:0")
    g = load_golden("module_d32")
    m, (C, M, L, P, bf) = _load_module(g, torch.bfloat16, dev)
    q, v, refp, sh, lsi, mask = _inputs(g, torch.bfloat16, dev, bf)
    out = m(query=q, value=v, key_padding_mask=mask, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
    params = {k: p.detach().double().cpu() for k, p in m.state_dict().items()}
    truth = O.module_forward(params, q.detach().double().cpu(), v.detach().double().cpu(), None, refp.double().cpu(),
                             sh.cpu(), M, L, P)
    scale = truth.abs().max().item()
    assert (out.detach().double().cpu() - truth).abs().max().item() < 1e-2 * scale      # north_star: bf16 within 1e-2


def test_module_bf16_scaled_fp16_accumulation_switch():
    """fused.f16_accumulate (opt-in, MSDA_B200_F16ACC=1): the module's input gradients with grad_value accumulated in scaled
    fp16 against the default fp32 accumulation -- output and grad_query identical (the query side does not depend on the
    map), grad_value_in within 1e-2 of its maximum (measured ~1e-3).  256 channels / 8 heads / 4 levels / 4 points: the
    shape class that takes the fused-query backward."""
    import ziragroundingdino_b200 as zb
    from ziragroundingdino_b200 import fused
    dev = torch.device("cuda:0")
    torch.manual_seed(12)
    C, M, L, P, N, Lq = 256, 8, 4, 4, 2, 700
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    m = zb.MultiScaleDeformableAttention(C, M, L, P, batch_first=True)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.01)
        m.attention_weights.weight.normal_(0, 0.05)
    m = m.to(dev).to(torch.bfloat16)
    sh = torch.tensor(shapes, device=dev)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    q0 = torch.randn(N, Lq, C, device=dev).to(torch.bfloat16)
    v0 = torch.randn(N, S, C, device=dev).to(torch.bfloat16)
    refp = torch.rand(N, Lq, L, 2, device=dev)
    gy = torch.randn(N, Lq, C, device=dev).to(torch.bfloat16)
    res = {}
    keep = fused.f16_accumulate
    try:
        for mode in (False, True):
            fused.f16_accumulate = mode
            q, v = q0.clone().requires_grad_(True), v0.clone().requires_grad_(True)
            out = m(query=q, value=v, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
            out.backward(gy)
            res[mode] = (out.detach().float(), q.grad.float(), v.grad.float())
    finally:
        fused.f16_accumulate = keep
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    d = (res[True][2] - res[False][2]).abs().max().item() / res[False][2].abs().max().item()
    print("module bf16: grad_value_in, scaled-fp16 vs fp32 accumulation: max diff / max = %.2e" % d)
    assert 0 < d < 1e-2


def test_module_errors():
    import ziragroundingdino_b200 as zb
    dev = torch.device("cuda:0")
    m = zb.MultiScaleDeformableAttention(64, 2, 2, 4, batch_first=True).to(dev)
    sh = torch.tensor([[3, 2], [1, 1]], device=dev)
    lsi = torch.tensor([0, 6], device=dev)
    x = torch.randn(1, 7, 64, device=dev)
    with pytest.raises(ValueError, match="Last dim of reference_points"):
        m(query=x, value=x, reference_points=torch.rand(1, 7, 2, 3, device=dev), spatial_shapes=sh, level_start_index=lsi)
    with pytest.raises(AssertionError):
        m(query=x, value=x[:, :6], reference_points=torch.rand(1, 7, 2, 2, device=dev), spatial_shapes=sh,
          level_start_index=lsi)
    # fp16 keeps working (the reference up-casts around the op, ms_deform_attn.py:326-344)
    mh = m.half()
    o = mh(query=x.half(), value=x.half(), reference_points=torch.rand(1, 7, 2, 2, device=dev).half(),
           spatial_shapes=sh, level_start_index=lsi)
    assert o.dtype == torch.float16 and torch.isfinite(o).all()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_zira_branches_train_eval_merge(dtype):
    """ZiRa on value_proj / output_proj: train-mode (unmerged) == after __rep__ (any mode) within 1e-5;
    eval-mode ignores the un-merged branch; the zero-inter loss equals the oracle's RepZeroLinear
    restatement composed with the base linear."""
    import ziragroundingdino_b200 as zb
    dev = torch.device("cuda:0")
    g = load_golden("module_dec")
    m, (C, M, L, P, bf) = _load_module(g, dtype, dev)
    m.add_zira_branches()
    torch.manual_seed(0)
    with torch.no_grad():
        for ad in (m.value_proj_adapter, m.output_proj_adapter):
            ad.weight.normal_(0, 1e-2); ad.bias.normal_(0, 1e-2)
            ad.freeze_linear.weight.normal_(0, 1e-2); ad.freeze_linear.bias.normal_(0, 1e-2)
    q, v, refp, sh, lsi, mask = _inputs(g, dtype, dev, bf)
    kw = dict(query=q, value=v, key_padding_mask=mask, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
    m.train()
    y_train = m(**kw)
    loss = m.zero_inter_loss
    assert loss is not None and loss.requires_grad
    (y_train.square().mean() + 0.1 * loss).backward()
    for n, p in m.named_parameters():
        if "adapter" in n:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # oracle: value_proj with the 3-term weight, loss from RepZeroLinear semantics
    ad = m.value_proj_adapter
    o_out, o_loss_v = O.rep_zero_linear(v.detach().cpu(), ad.weight.detach().cpu(), ad.bias.detach().cpu(),
                                        ad.scaling.detach().cpu(), ad.freeze_linear.weight.detach().cpu(),
                                        ad.freeze_linear.bias.detach().cpu(), training=True)
    m.eval()
    y_eval_unmerged = m(**kw)
    assert m.zero_inter_loss is None
    assert (y_eval_unmerged - y_train).abs().max() > 1e-6      # branch really contributes in train mode
    zb.merge_all(m)
    y_eval_merged = m(**kw)
    m.train()
    y_train_merged = m(**kw)
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert (y_eval_merged - y_train).abs().max().item() < tol
    assert (y_train_merged - y_train).abs().max().item() < max(tol, 1e-6)
    assert float(m.value_proj_adapter.scaling) == pytest.approx(0.1)
    assert float(m.value_proj_adapter.weight.abs().max()) == pytest.approx(1e-8)
    assert float(o_loss_v) > 0


def test_module_under_autocast():
    """fp32 module under torch.autocast(bf16) (the reference trainer's AMP switch, train_multidatasets.py:169): the
    projections autocast to bf16, the op runs its 16-bit kernels with fp32 locations/weights; result close to fp32."""
    import ziragroundingdino_b200 as zb
    dev = torch.device("cuda:0")
    g = load_golden("module_d32")
    m, (C, M, L, P, bf) = _load_module(g, torch.float32, dev)
    q, v, refp, sh, lsi, mask = _inputs(g, torch.float32, dev, bf)
    kw = dict(key_padding_mask=mask, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
    y32 = m(query=q, value=v, **kw)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = m(query=q, value=v, **kw)
    assert y16.dtype == torch.bfloat16
    assert (y16.float() - y32).abs().max().item() < 3e-2 * y32.abs().max().item()
    y16.float().square().mean().backward()
    assert torch.isfinite(q.grad).all() and torch.isfinite(v.grad).all()


def test_module_many_images_and_single_query():
    """Shapes at the edges: a batch larger than im2col_step's default divisor logic cares about, and Lq = 1."""
    import ziragroundingdino_b200 as zb
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = zb.MultiScaleDeformableAttention(256, 8, 4, 4, batch_first=True).to(dev)
    shapes = [(4, 5), (2, 3), (1, 2), (1, 1)]
    S = sum(h * w for h, w in shapes)
    sh = torch.tensor(shapes, device=dev)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    for N, Lq in ((64, 3), (128, 1), (96, 2)):          # 96 % min(96, 64) != 0 -> the reference's observable error
        x = torch.randn(N, S, 256, device=dev)
        q = torch.randn(N, Lq, 256, device=dev)
        refp = torch.rand(N, Lq, 4, 2, device=dev)
        call = lambda: m(query=q, value=x, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
        if N % min(N, 64):
            with pytest.raises(RuntimeError, match="must divide im2col_step"):
                call()
        else:
            y = call()
            assert y.shape == (N, Lq, 256) and torch.isfinite(y).all()
