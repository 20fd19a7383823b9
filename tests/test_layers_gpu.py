"""GPU: row N1 of SURVEY.md 8(f) -- the encoder / decoder layers around MultiScaleDeformableAttention -- against
fixtures generated from the REFERENCE's own layer classes (tests/golden/make_golden.py: layer_cases, fp64, CPU).
fp64 on the GPU runs the generic kernels + library linears: bar 1e-9 (same arithmetic, different summation order).
bf16 runs the fused path (tcgen05 GEMMs, fused add+LayerNorm / FFN): outputs are LayerNorm-normalised O(1) values, bar
5e-2 absolute on outputs (a few bf16 roundings through 3-4 normalisations), 8e-2 relative (Frobenius) on input gradients."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load(layer, g, dtype):
    sd = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    layer = layer.double()
    layer.load_state_dict(sd)          # strict: the reference's key names must match ours exactly
    return layer.to(DEV).to(dtype).eval()


def _t(g, k, dtype=None):
    t = torch.from_numpy(g[k]).to(DEV)
    return t.to(dtype) if dtype is not None and t.is_floating_point() else t


@pytest.mark.parametrize("dtype,out_tol,grad_tol", [(torch.float64, 1e-9, 1e-8), (torch.float32, 2e-4, 2e-3)])
def test_encoder_layer_vs_reference_fixture(dtype, out_tol, grad_tol):
    import ziragroundingdino_b200 as zb
    g = load_golden("layer_encoder")
    C, FF, M, L, P = (int(v) for v in g["cfg"])
    layer = _load(zb.DeformableTransformerEncoderLayer(C, FF, 0.0, "relu", L, M, P, use_adapter=False), g, dtype)
    sh = _t(g, "shapes")
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    src = _t(g, "src", dtype).requires_grad_(True)
    y, aloss = layer(src, _t(g, "pos", dtype), _t(g, "reference_points", dtype), sh, lsi, _t(g, "mask"))
    y.backward(_t(g, "grad_out", dtype))
    assert float(aloss) == 0.0
    assert np.abs(y.detach().double().cpu().numpy() - g["out"]).max() < out_tol
    assert rel_err(src.grad.double().cpu(), torch.from_numpy(g["grad_src"])) < grad_tol


@pytest.mark.parametrize("dtype,out_tol,grad_tol", [(torch.float64, 1e-9, 1e-8), (torch.float32, 2e-4, 2e-3)])
def test_decoder_layer_vs_reference_fixture(dtype, out_tol, grad_tol):
    import ziragroundingdino_b200 as zb
    g = load_golden("layer_decoder")
    C, FF, M, L, P = (int(v) for v in g["cfg"])
    layer = _load(zb.DeformableTransformerDecoderLayer(C, FF, 0.0, "relu", L, M, P, use_text_cross_attention=True,
                                                       use_adapter=False), g, dtype)
    sh = _t(g, "shapes")
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    tgt = _t(g, "tgt", dtype).requires_grad_(True)
    mem = _t(g, "memory", dtype).requires_grad_(True)
    y, aloss = layer(tgt=tgt, tgt_query_pos=_t(g, "query_pos", dtype), tgt_reference_points=_t(g, "reference_points", dtype),
                     memory_text=_t(g, "memory_text", dtype), text_attention_mask=_t(g, "text_mask"), memory=mem,
                     memory_key_padding_mask=_t(g, "mask"), memory_level_start_index=lsi, memory_spatial_shapes=sh)
    y.backward(_t(g, "grad_out", dtype))
    assert float(aloss) == 0.0
    assert np.abs(y.detach().double().cpu().numpy() - g["out"]).max() < out_tol
    assert rel_err(tgt.grad.double().cpu(), torch.from_numpy(g["grad_tgt"])) < grad_tol
    assert rel_err(mem.grad.double().cpu(), torch.from_numpy(g["grad_memory"])) < grad_tol


def test_layers_reject_the_bottleneck_adapter():
    import ziragroundingdino_b200 as zb
    with pytest.raises(NotImplementedError):
        zb.DeformableTransformerEncoderLayer(use_adapter=True)
    with pytest.raises(NotImplementedError):
        zb.DeformableTransformerDecoderLayer(use_adapter=True)


def test_decoder_layer_bf16_fused_vs_oracle():
    """Model-sized decoder layer (C=256, 900 queries, Swin-T memory) in bf16 on the fused path against the oracle
    restatement (oracle/cpu_encoder.decoder_layer, pinned to the reference fixture) in fp64 on the same parameters."""
    import ziragroundingdino_b200 as zb
    from oracle import cpu_encoder
    torch.manual_seed(3)
    C, FF, M, L, P, N, nq, nt = 256, 2048, 8, 4, 4, 2, 300, 16
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    layer = zb.DeformableTransformerDecoderLayer(C, FF, 0.0, "relu", L, M, P, use_text_cross_attention=True)
    with torch.no_grad():
        layer.cross_attn.sampling_offsets.weight.normal_(0, 0.01)
        layer.cross_attn.attention_weights.weight.normal_(0, 0.05)
    layer = layer.to(DEV).to(torch.bfloat16).eval()
    sh = torch.tensor(shapes, device=DEV)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    bf = lambda *s: torch.randn(*s, device=DEV).to(torch.bfloat16)
    tgt, qpos, mem, text = bf(nq, N, C), bf(nq, N, C), bf(S, N, C), bf(N, nt, C)
    ref4 = torch.cat([torch.rand(nq, N, L, 2, device=DEV) * 0.8 + 0.1, torch.rand(nq, N, L, 2, device=DEV) * 0.4 + 0.05], -1).to(torch.bfloat16)
    mask = torch.zeros(N, S, dtype=torch.bool, device=DEV)
    mask[1, S - 100:] = True
    tmask = torch.zeros(N, nt, dtype=torch.bool, device=DEV)
    tmask[0, -3:] = True
    y, _ = layer(tgt=tgt, tgt_query_pos=qpos, tgt_reference_points=ref4, memory_text=text, text_attention_mask=tmask,
                 memory=mem, memory_key_padding_mask=mask, memory_level_start_index=lsi, memory_spatial_shapes=sh)
    d = lambda t: t.detach().double().cpu()
    p = {k: d(v) for k, v in layer.state_dict().items()}
    ref = cpu_encoder.decoder_layer(p, d(tgt), d(qpos), d(ref4), d(mem), d(text), tmask.cpu(), sh.cpu(), mask.cpu(), M, L, P)
    assert (d(y) - ref).abs().max().item() < 5e-2
    assert rel_err(d(y), ref) < 1e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("with_mask", [True, False])
def test_encoder_layer_block_functions_vs_stagewise(dtype, with_mask):
    """Frozen encoder layer (the ZiRa configuration): each half as ONE autograd Function whose dgrad GEMMs accumulate onto
    the residual gradient (blocks.py, msda_linear_accum_16) vs the stage-wise Functions + autograd's elementwise adds.
    Forward runs the same kernels (bit-equal); the input gradient differs only by where the 16-bit roundings fall, and
    both are checked against an fp64 evaluation of the layer (oracle.cpu_encoder.encoder_layer) on the same parameters."""
    import ziragroundingdino_b200 as zb
    from oracle import cpu_encoder
    torch.manual_seed(5)
    C, FF, M, L, P, N = 256, 2048, 8, 4, 4, 2
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    layer = zb.DeformableTransformerEncoderLayer(C, FF, 0.0, "relu", L, M, P)
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
        layer.norm1.weight.normal_(1, 0.1); layer.norm2.bias.normal_(0, 0.1)
    layer = layer.to(DEV).to(dtype)
    for p in layer.parameters():
        p.requires_grad_(False)
    sh = torch.tensor(shapes, device=DEV)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    src0 = torch.randn(N, S, C, device=DEV).to(dtype)
    pos = torch.randn(N, S, C, device=DEV).to(dtype)
    refp = torch.rand(N, S, L, 2, device=DEV)
    mask = None
    if with_mask:
        mask = torch.zeros(N, S, dtype=torch.bool, device=DEV)
        mask[1, S - 150:] = True
    gy = torch.randn(N, S, C, device=DEV).to(dtype)
    res = {}
    for blocks_on in (True, False):
        zb.DeformableTransformerEncoderLayer.block_functions = blocks_on
        try:
            x = src0.clone().requires_grad_(True)
            y, _ = layer(x, pos, refp, sh, lsi, mask)
            y.backward(gy)
            res[blocks_on] = (y.detach(), x.grad)
        finally:
            zb.DeformableTransformerEncoderLayer.block_functions = True
    # the block form keeps x + ffn(x) in fp32 until LayerNorm's input is rounded (the stage-wise path rounds ffn(x) first)
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert (res[True][0].float() - res[False][0].float()).abs().max().item() <= 4 * ulp * res[False][0].float().abs().max().item()
    # the block form also projects src + pos WITHOUT rounding the sum to 16 bits (blocks.FOLD_POS): a few sampling
    # locations land on the other side of a pixel boundary than in the stage-wise path, whose query is the rounded sum
    assert rel_err(res[True][1].float().cpu(), res[False][1].float().cpu()) < 4e-2
    # fp64 truth of the same layer
    d = lambda t: t.detach().double().cpu()
    p = {(k[len("self_attn."):] if k.startswith("self_attn.") else k): d(v) for k, v in layer.state_dict().items()}
    xd = d(src0).requires_grad_(True)
    ref = cpu_encoder.encoder_layer(p, xd, d(pos), d(refp), sh.cpu(), None if mask is None else mask.cpu(), M, L, P)
    ref.backward(d(gy))
    assert (d(res[True][0]) - ref.detach()).abs().max().item() < 6e-2
    e_block, e_stage = rel_err(d(res[True][1]), xd.grad), rel_err(d(res[False][1]), xd.grad)
    # bf16 carries 8 mantissa bits through ~10 rounded intermediates; the bar is the stage-wise path's own distance
    assert e_block < (6e-2 if dtype == torch.bfloat16 else 3e-2) and e_block <= e_stage * 1.2 + 1e-3


def test_linear_accum16_in_place_and_out_of_place():
    from ziragroundingdino_b200 import blocks, _lib
    for dtype, eps in ((torch.bfloat16, 2 ** -8), (torch.float16, 2 ** -11)):
        for R, K, Nout in ((1000, 256, 256), (4 * 22223 + 3, 384, 256), (77, 512, 64)):
            g = torch.Generator().manual_seed(R)
            x = torch.randn(R, K, generator=g).to(dtype).to(DEV)
            w = (torch.randn(Nout, K, generator=g) * 0.06).to(dtype).to(DEV)
            acc = torch.randn(R, Nout, generator=g).to(dtype).to(DEV)
            want = acc.double() + x.double() @ w.double().t()
            out = torch.empty_like(acc)
            for staged in (1, 0):
                _lib.lib().msda_b200_gemm_set_staged(staged)
                try:
                    blocks.linear_accum16(x, w, acc, out)
                    a2 = acc.clone()
                    blocks.linear_accum16(x, w, a2)
                finally:
                    _lib.lib().msda_b200_gemm_set_staged(1)
                assert (out.double() - want).abs().max().item() <= eps * want.abs().max().item() * 1.05
                assert torch.equal(out, a2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R", [1, 127, 128, 1000, 22223])
def test_ffn_chain_forward_and_backward_vs_fp64(dtype, R):
    """Chained FFN kernel (hidden activation never leaves the SM) vs fp64 on the same 16-bit operands, with the SAME
    intermediate rounding the two-launch path has (hidden rounded to the storage type before the second product):
    forward y = relu(x W1^T + b1) W2^T + b2, the 1-bit ReLU mask, and backward dx = dz + ((dz W2) * mask) W1 in place."""
    from ziragroundingdino_b200 import blocks
    dev = "cuda:0"
    g = torch.Generator().manual_seed(5 + R)
    C, F = 256, 2048
    x = torch.randn(R, C, generator=g).to(dtype).to(dev)
    w1 = (torch.randn(F, C, generator=g) * 0.06).to(dtype).to(dev)
    w2 = (torch.randn(C, F, generator=g) * 0.02).to(dtype).to(dev)
    b1 = (torch.randn(F, generator=g) * 0.1).to(dev)
    b2 = (torch.randn(C, generator=g) * 0.1).to(dev)
    bits = torch.zeros((F // 32, R), dtype=torch.int32, device=dev)
    y = blocks.ffn_chain_fwd16(x, w1, b1, w2, b2, bits)
    h64 = torch.relu(x.double() @ w1.double().t() + b1.double())
    h16 = h64.to(dtype).double()
    want = h16 @ w2.double().t() + b2.double()
    eps = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert (y.double() - want).abs().max().item() <= 2 * eps * want.abs().max().item()
    # bits: exact wherever the pre-activation is not within rounding distance of zero
    pre = x.double() @ w1.double().t() + b1.double()
    got = ((bits.t().contiguous().view(R, F // 32, 1) >> torch.arange(32, device=dev).view(1, 1, 32)) & 1).reshape(R, F).bool()
    sure = pre.abs() > 1e-3
    assert torch.equal(got[sure], (pre > 0)[sure])
    dz = torch.randn(R, C, generator=g).to(dtype).to(dev)
    dz_in = dz.clone()
    dx = blocks.ffn_chain_bwd16(dz, w2.t().contiguous(), w1.t().contiguous(), bits)
    assert dx.data_ptr() == dz.data_ptr()
    dh = ((dz_in.double() @ w2.double()) * got.double()).to(dtype).double()
    want_dx = dz_in.double() + dh @ w1.double()
    assert (dx.double() - want_dx).abs().max().item() <= 2 * eps * want_dx.abs().max().item()


def test_encoder_layer_with_chained_ffn_matches_unchained():
    """The frozen encoder layer with blocks.FFN_CHAIN on vs off: same outputs and input gradients to 16-bit rounding."""
    import ziragroundingdino_b200 as zb
    from ziragroundingdino_b200 import blocks, encoder, synthetic as syn
    dev = "cuda:0"
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    sh, lsi = syn.level_tensors(shapes, dev)
    torch.manual_seed(9)
    layer = encoder.DeformableTransformerEncoderLayer(256, 2048, 0.0, "relu", 4, 8, 4)
    with torch.no_grad():      # (the gradient of mean(out^2) through a LayerNorm with gamma = 1, beta = 0 is exactly zero: use
        for nrm in (layer.norm1, layer.norm2):      # a random affine and a random upstream gradient instead of rounding noise)
            nrm.weight.normal_(1, 0.2); nrm.bias.normal_(0, 0.2)
    layer = layer.to(dev).bfloat16()
    for p in layer.parameters():
        p.requires_grad_(False)
    refp = syn.encoder_reference_points(shapes, torch.ones(2, 4, 2, device=dev), dev)
    x0 = torch.randn(2, S, 256, device=dev).bfloat16()
    pos = torch.randn(2, S, 256, device=dev).bfloat16()
    gout = torch.randn(2, S, 256, device=dev).bfloat16()
    res = {}
    keep = blocks.FFN_CHAIN
    try:
        for chain in (False, True):
            blocks.FFN_CHAIN = chain
            x = x0.clone().requires_grad_(True)
            out, _ = layer(x, pos, refp, sh, lsi, None)
            out.backward(gout)
            res[chain] = (out.detach().float(), x.grad.float())
    finally:
        blocks.FFN_CHAIN = keep
    assert (res[True][0] - res[False][0]).abs().max().item() < 3 * 2 ** -8 * res[False][0].abs().max().item()
    assert (res[True][1] - res[False][1]).abs().max().item() < 2e-2 * res[False][1].abs().max().item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R", [1, 127, 128, 1000, 22223])
def test_ffn_chain_with_fused_residual_layernorm_vs_fp64(dtype, R):
    """The chained FFN with `norm2(x + ffn(x))` in its final stage: z (LayerNorm's saved input, 16 bit), y, mean, rstd against
    fp64 on the same operands (hidden activation rounded to the storage type, as in the kernel)."""
    from ziragroundingdino_b200 import blocks
    dev = "cuda:0"
    g = torch.Generator().manual_seed(11 + R)
    C, F = 256, 2048
    x = torch.randn(R, C, generator=g).to(dtype).to(dev)
    w1 = (torch.randn(F, C, generator=g) * 0.06).to(dtype).to(dev)
    w2 = (torch.randn(C, F, generator=g) * 0.02).to(dtype).to(dev)
    b1 = (torch.randn(F, generator=g) * 0.1).to(dev)
    b2 = (torch.randn(C, generator=g) * 0.1).to(dev)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    bits = torch.zeros((F // 32, R), dtype=torch.int32, device=dev)
    z, y, mean, rstd = blocks.ffn_chain_ln_fwd16(x, w1, b1, w2, b2, gamma, beta, 1e-5, bits)
    h16 = torch.relu(x.double() @ w1.double().t() + b1.double()).to(dtype).double()
    z64 = x.double() + h16 @ w2.double().t() + b2.double()
    eps = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert (z.double() - z64).abs().max().item() <= 2 * eps * z64.abs().max().item()
    zr = z.double()                       # LayerNorm is defined on the ROUNDED sum
    m64, v64 = zr.mean(1), zr.var(1, unbiased=False)
    assert (mean.double() - m64).abs().max().item() < 1e-5 * max(1.0, m64.abs().max().item())
    assert ((rstd.double() - (v64 + 1e-5).rsqrt()).abs() * (v64 + 1e-5).sqrt()).max().item() < 1e-4
    y64 = (zr - m64[:, None]) * (v64 + 1e-5).rsqrt()[:, None] * gamma.double() + beta.double()
    assert (y.double() - y64).abs().max().item() <= 2 * eps * y64.abs().max().item()
    # same ReLU mask as the plain chain
    bits2 = torch.zeros_like(bits)
    blocks.ffn_chain_fwd16(x, w1, b1, w2, b2, bits2)
    assert torch.equal(bits, bits2)


def test_encoder_layer_with_fused_ffn_layernorm_matches_separate():
    """The frozen encoder layer with blocks.FFN_LN on vs off: same outputs and input gradients to 16-bit rounding."""
    from ziragroundingdino_b200 import blocks, encoder, synthetic as syn
    dev = "cuda:0"
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    sh, lsi = syn.level_tensors(shapes, dev)
    torch.manual_seed(10)
    layer = encoder.DeformableTransformerEncoderLayer(256, 2048, 0.0, "relu", 4, 8, 4)
    with torch.no_grad():      # (the gradient of mean(out^2) through a LayerNorm with gamma = 1, beta = 0 is exactly zero: use
        for nrm in (layer.norm1, layer.norm2):      # a random affine and a random upstream gradient instead of rounding noise)
            nrm.weight.normal_(1, 0.2); nrm.bias.normal_(0, 0.2)
    layer = layer.to(dev).bfloat16()
    for p in layer.parameters():
        p.requires_grad_(False)
    refp = syn.encoder_reference_points(shapes, torch.ones(2, 4, 2, device=dev), dev)
    x0 = torch.randn(2, S, 256, device=dev).bfloat16()
    pos = torch.randn(2, S, 256, device=dev).bfloat16()
    gout = torch.randn(2, S, 256, device=dev).bfloat16()
    res = {}
    keep = blocks.FFN_LN
    try:
        for fused_ln in (False, True):
            blocks.FFN_LN = fused_ln
            x = x0.clone().requires_grad_(True)
            out, _ = layer(x, pos, refp, sh, lsi, None)
            out.backward(gout)
            res[fused_ln] = (out.detach().float(), x.grad.float())
    finally:
        blocks.FFN_LN = keep
    assert (res[True][0] - res[False][0]).abs().max().item() < 3 * 2 ** -8 * res[False][0].abs().max().item()
    assert (res[True][1] - res[False][1]).abs().max().item() < 2e-2 * res[False][1].abs().max().item()


def test_ffn_chain_pair_kernel_bitwise_equals_single_cta_kernel():
    """The CTA-pair chained FFN (tcgen05 cta_group::2, cross-CTA barriers with relaxed arrivals) must reproduce the single-CTA
    kernel bit for bit -- same products, same accumulation order, same epilogue arithmetic (the LayerNorm statistics add four
    partial sums per row instead of two) -- and itself on every repetition (a race in the cross-CTA signalling would show up
    as a sporadic mismatch)."""
    from ziragroundingdino_b200 import blocks
    dev = "cuda:0"
    g = torch.Generator().manual_seed(77)
    C, F = 256, 2048
    keep = blocks.FFN_PAIR
    try:
        for R in (88892, 1000, 129):
            x = torch.randn(R, C, generator=g).bfloat16().to(dev)
            w1 = (torch.randn(F, C, generator=g) * 0.06).bfloat16().to(dev)
            w2 = (torch.randn(C, F, generator=g) * 0.02).bfloat16().to(dev)
            b1 = (torch.randn(F, generator=g) * 0.1).to(dev)
            b2 = (torch.randn(C, generator=g) * 0.1).to(dev)
            gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
            beta = (0.1 * torch.randn(C, generator=g)).to(dev)
            dz0 = torch.randn(R, C, generator=g).bfloat16().to(dev)
            w1t, w2t = w1.t().contiguous(), w2.t().contiguous()

            def run():
                bits = torch.zeros((F // 32, R), dtype=torch.int32, device=dev)
                y = blocks.ffn_chain_fwd16(x, w1, b1, w2, b2, bits)
                bits2 = torch.zeros_like(bits)
                z, yl, mean, rstd = blocks.ffn_chain_ln_fwd16(x, w1, b1, w2, b2, gamma, beta, 1e-5, bits2)
                dx = blocks.ffn_chain_bwd16(dz0.clone(), w2t, w1t, bits)
                return y, bits, z, yl, mean, rstd, bits2, dx
            blocks.FFN_PAIR = False
            ref = run()
            blocks.FFN_PAIR = True
            first = run()
            names = ("y", "bits", "z", "y_ln", "mean", "rstd", "bits_ln", "dx")
            for n, a, b in zip(names, ref, first):
                if n in ("y_ln", "mean", "rstd"):      # four partial sums per row instead of two: last-bit differences allowed
                    tol = 2 ** -7 if n == "y_ln" else 1e-5
                    assert (a.double() - b.double()).abs().max().item() <= tol * max(1.0, a.double().abs().max().item()), n
                else:
                    assert torch.equal(a, b), n
            for _ in range(12 if R > 10000 else 40):
                got = run()
                for n, a, b in zip(names, first, got):
                    assert torch.equal(a, b), n
    finally:
        blocks.FFN_PAIR = keep


def test_linear_accum2_two_operands_k_concatenated():
    """out = accum + [x1 | x2] w^T with the k-blocks of ONE GEMM coming from two activation tensors (msda_linear_accum2_16):
    against fp64, against the product over a materialised concatenation (bit-equal: same tiles, same order), in place
    and out of place, rows not a multiple of the tile, resident and streamed W."""
    from ziragroundingdino_b200 import blocks, _lib
    for dtype, eps in ((torch.bfloat16, 2 ** -8), (torch.float16, 2 ** -11)):
        for R, K1, K2, Nout in ((1000, 256, 384, 256), (4 * 22223 + 3, 256, 384, 256), (77, 64, 64, 64), (300, 128, 512, 128)):
            g = torch.Generator().manual_seed(R + K2)
            x1 = torch.randn(R, K1, generator=g).to(dtype).to(DEV)
            x2 = torch.randn(R, K2, generator=g).to(dtype).to(DEV)
            w = (torch.randn(Nout, K1 + K2, generator=g) * 0.05).to(dtype).to(DEV)
            acc = torch.randn(R, Nout, generator=g).to(dtype).to(DEV)
            want = acc.double() + torch.cat([x1, x2], 1).double() @ w.double().t()
            out = torch.empty_like(acc)
            for resident in (1, 0):
                _lib.lib().msda_b200_gemm_set_resident(resident)
                try:
                    blocks.linear_accum2_16(x1, x2, w, acc, out)
                    a2 = acc.clone()
                    blocks.linear_accum2_16(x1, x2, w, a2)
                    cat = blocks.linear_accum16(torch.cat([x1, x2], 1).contiguous(), w, acc, torch.empty_like(acc))
                finally:
                    _lib.lib().msda_b200_gemm_set_resident(1)
                assert (out.double() - want).abs().max().item() <= eps * want.abs().max().item() * 1.05
                assert torch.equal(out, a2) and torch.equal(out, cat)


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_query_projection_of_a_sum_without_forming_it(dtype, ref_dim):
    """msda_query_proj2_16(src, pos) = sampling locations / attention weights of (src + pos), the two operands streamed
    against one resident copy of the stacked weight.  Truth: fp64 on the exact sum; the two-operand form must be at least
    as close as projecting the rounded 16-bit sum (it skips that rounding)."""
    from ziragroundingdino_b200 import fused
    torch.manual_seed(3)
    C, M, L, P, R = 256, 8, 4, 4, 4 * 1234 + 5
    shapes = torch.tensor([(20, 30), (10, 15), (5, 8), (3, 4)], device=DEV)
    src = torch.randn(R, C, device=DEV).to(dtype)
    pos = torch.randn(R, C, device=DEV).to(dtype)
    w_cat = (torch.randn(3 * M * L * P, C, device=DEV) * 0.03).to(dtype)
    b_cat = torch.randn(3 * M * L * P, device=DEV) * 0.1
    ref = torch.rand(R, L, ref_dim, device=DEV) * 0.6 + 0.2
    loc2, aw2 = fused.query_proj16(src, w_cat, b_cat, ref, ref_dim, shapes, M, L, P, q_add=pos)
    loc1, aw1 = fused.query_proj16((src + pos).contiguous(), w_cat, b_cat, ref, ref_dim, shapes, M, L, P)
    raw = (src.double() + pos.double()) @ w_cat.double().t() + b_cat.double()
    off = raw[:, :2 * M * L * P].view(R, M, L, P, 2)
    awt = torch.softmax(raw[:, 2 * M * L * P:].view(R, M, L * P), -1).view(R, M, L, P)
    if ref_dim == 2:
        norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).double()
        loct = ref.double()[:, None, :, None, :] + off / norm[None, None, :, None, :]
    else:
        loct = ref.double()[:, None, :, None, :2] + off / P * ref.double()[:, None, :, None, 2:] * 0.5
    e2l, e1l = (loc2.double() - loct).abs().max().item(), (loc1.double() - loct).abs().max().item()
    e2a, e1a = (aw2.double() - awt).abs().max().item(), (aw1.double() - awt).abs().max().item()
    print("query projection of src + pos (%s, ref_dim %d): loc err two-operand %.2e / rounded sum %.2e; aw %.2e / %.2e" % (dtype, ref_dim, e2l, e1l, e2a, e1a))
    assert e2l <= e1l * 1.05 + 1e-7 and e2a <= e1a * 1.05 + 1e-7
    assert e2l < 1e-4 and e2a < 1e-5          # fp32 accumulation of exact 16-bit operands


def test_encoder_block_operand_folds_match_the_separate_launches():
    """blocks.FOLD_POS / blocks.DGRAD_CAT (default on): the encoder's attention block with `src + pos` folded into the query
    projection and its two input dgrads as one K-concatenated product, against the separate elementwise add and the two
    accumulating GEMMs.  The fold projects the exact sum (locations 7e-7 from fp64 instead of 2e-3 for the rounded 16-bit
    sum, see the test above), so some samples move across pixel boundaries: outputs agree to a few 16-bit ulps of the
    LayerNorm-scale values, the input gradient to 4e-2 (Frobenius) -- the distance the stage-wise comparison also shows."""
    import ziragroundingdino_b200 as zb
    from ziragroundingdino_b200 import blocks
    torch.manual_seed(9)
    C, FF, M, L, P, N = 256, 2048, 8, 4, 4, 2
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    layer = zb.DeformableTransformerEncoderLayer(C, FF, 0.0, "relu", L, M, P)
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    layer = layer.to(DEV).to(torch.bfloat16)
    for p in layer.parameters():
        p.requires_grad_(False)
    sh = torch.tensor(shapes, device=DEV)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    src0 = torch.randn(N, S, C, device=DEV).to(torch.bfloat16)
    pos = torch.randn(N, S, C, device=DEV).to(torch.bfloat16)
    refp = torch.rand(N, S, L, 2, device=DEV)
    mask = torch.zeros(N, S, dtype=torch.bool, device=DEV)
    mask[1, S - 150:] = True
    gy = torch.randn(N, S, C, device=DEV).to(torch.bfloat16)
    res = {}
    keep = (blocks.FOLD_POS, blocks.DGRAD_CAT)
    try:
        for on in (True, False):
            blocks.FOLD_POS = blocks.DGRAD_CAT = on
            x = src0.clone().requires_grad_(True)
            y, _ = layer(x, pos, refp, sh, lsi, mask)
            y.backward(gy)
            res[on] = (y.detach().float(), x.grad.float())
    finally:
        blocks.FOLD_POS, blocks.DGRAD_CAT = keep
    dy = (res[True][0] - res[False][0]).abs().max().item()
    dg = rel_err(res[True][1].cpu(), res[False][1].cpu())
    print("encoder block folds on vs off: max |dy| %.2e, rel d(src) %.2e" % (dy, dg))
    assert dy <= 6e-2 and dg < 4e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R,K,Nout", [(1, 256, 256), (127, 256, 256), (1000, 256, 256), (4 * 22223 + 3, 256, 256), (300, 128, 128), (777, 512, 256)])
def test_projection_residual_layernorm_in_one_launch(dtype, R, K, Nout):
    """msda_linear_add_layernorm_16 (z = residual + x W^T + b stored, y = LayerNorm(z), mean, rstd) against fp64 on the same
    16-bit operands, and against the two-launch form (msda_linear_16 + msda_add_layernorm_fwd_16): z within one 16-bit ulp
    (the fused form does not round the projection before the add), y within a few ulps of O(1) values, statistics consistent
    with the z it stored (the backward kernel reads exactly these)."""
    from ziragroundingdino_b200 import blocks, fused
    g = torch.Generator().manual_seed(R + K)
    x = torch.randn(R, K, generator=g).to(dtype).to(DEV)
    w = (torch.randn(Nout, K, generator=g) * 0.06).to(dtype).to(DEV)
    bias = (torch.randn(Nout, generator=g) * 0.1).to(DEV)
    res = torch.randn(R, Nout, generator=g).to(dtype).to(DEV)
    gamma = (1 + 0.1 * torch.randn(Nout, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(Nout, generator=g)).to(DEV)
    eps = 1e-5
    z, y, mean, rstd = blocks.linear_add_ln16(x, w, bias, res, gamma, beta, eps)
    attn = fused.linear16(x, w, bias)
    z0, y0, mean0, rstd0 = blocks._add_ln_fwd(res, attn, gamma, beta, eps)
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    zt = res.double() + x.double() @ w.double().t() + bias.double()
    zmax = zt.abs().max().item()
    assert (z.double() - zt).abs().max().item() <= 1.01 * ulp * zmax
    assert (z.float() - z0.float()).abs().max().item() <= 2 * ulp * zmax
    # statistics of the stored z
    zs = z.double()
    m_ref, v_ref = zs.mean(1), zs.var(1, unbiased=False)
    assert (mean.double() - m_ref).abs().max().item() < 1e-4
    assert ((rstd.double() - (v_ref + eps).rsqrt()).abs() / (v_ref + eps).rsqrt()).max().item() < 1e-4
    yt = (zs - m_ref[:, None]) * (v_ref[:, None] + eps).rsqrt() * gamma.double() + beta.double()
    assert (y.double() - yt).abs().max().item() <= 2.5 * ulp * yt.abs().max().item()
    assert (y.float() - y0.float()).abs().max().item() <= 8 * ulp * yt.abs().max().item()


def test_encoder_block_with_layernorm_in_the_projection_matches_separate():
    """blocks.OUT_LN (default on) against the projection + add+LayerNorm pair, forward output and input gradient of the
    encoder's attention block."""
    import ziragroundingdino_b200 as zb
    from ziragroundingdino_b200 import blocks
    torch.manual_seed(10)
    C, FF, M, L, P, N = 256, 2048, 8, 4, 4, 2
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    layer = zb.DeformableTransformerEncoderLayer(C, FF, 0.0, "relu", L, M, P)
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
        layer.self_attn.attention_weights.weight.normal_(0, 0.05)
        layer.norm1.weight.normal_(1, 0.1); layer.norm1.bias.normal_(0, 0.1)
    layer = layer.to(DEV).to(torch.bfloat16)
    for p in layer.parameters():
        p.requires_grad_(False)
    sh = torch.tensor(shapes, device=DEV)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    src0 = torch.randn(N, S, C, device=DEV).to(torch.bfloat16)
    pos = torch.randn(N, S, C, device=DEV).to(torch.bfloat16)
    refp = torch.rand(N, S, L, 2, device=DEV)
    gy = torch.randn(N, S, C, device=DEV).to(torch.bfloat16)
    res = {}
    keep = blocks.OUT_LN
    try:
        for on in (True, False):
            blocks.OUT_LN = on
            x = src0.clone().requires_grad_(True)
            y, _ = layer(x, pos, refp, sh, lsi, None)
            y.backward(gy)
            res[on] = (y.detach().float(), x.grad.float())
    finally:
        blocks.OUT_LN = keep
    dy = (res[True][0] - res[False][0]).abs().max().item()
    dg = rel_err(res[True][1].cpu(), res[False][1].cpu())
    print("encoder block, LayerNorm in the projection on vs off: max |dy| %.2e, rel d(src) %.2e" % (dy, dg))
    assert dy <= 6e-2 and dg < 2e-2


@pytest.mark.parametrize("offset", [8.0, 50.0])
def test_layernorm_epilogues_with_a_common_offset(offset):
    """The fused LayerNorm epilogues (output projection, chained FFN) take the variance as E[z^2] - mean^2 in fp32 from ONE
    pass over the row.  That form loses precision when |mean| >> std; this pins how much for rows that all carry a common
    offset of 8 and of 50 standard deviations (LayerNorm inputs of the layers have |mean| / std < 1): rstd within 1e-3 of
    the two-pass value of the stored z, y within a few 16-bit ulps."""
    from ziragroundingdino_b200 import blocks
    g = torch.Generator().manual_seed(7)
    R, K, Nout = 1000, 256, 256
    x = torch.randn(R, K, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(Nout, K, generator=g) * 0.06).to(torch.bfloat16).to(DEV)
    res = (torch.randn(R, Nout, generator=g) + offset).to(torch.bfloat16).to(DEV)
    gamma, beta = torch.ones(Nout, device=DEV), torch.zeros(Nout, device=DEV)
    z, y, mean, rstd = blocks.linear_add_ln16(x, w, None, res, gamma, beta, 1e-5)
    zs = z.double()
    m_ref, r_ref = zs.mean(1), (zs.var(1, unbiased=False) + 1e-5).rsqrt()
    e_r = ((rstd.double() - r_ref).abs() / r_ref).max().item()
    yt = (zs - m_ref[:, None]) * r_ref[:, None]
    e_y = (y.double() - yt).abs().max().item()
    print("LayerNorm epilogue, common offset %g: rstd rel err %.2e, y abs err %.2e" % (offset, e_r, e_y))
    assert e_r < 1e-3 and e_y < 4 * 2 ** -8 * yt.abs().max().item()
