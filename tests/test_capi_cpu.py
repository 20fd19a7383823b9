"""CPU: the C-ABI library loads and exports every symbol include/msda_b200.h declares; argument
validation that needs no GPU; host-side module API parity with the reference (constructor, attribute
and state_dict names, init values, error behaviour)."""
import ctypes
import subprocess

import numpy as np
import pytest
import torch

import ziragroundingdino_b200 as zb
from conftest import load_golden
from ziragroundingdino_b200 import _lib


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert L.msda_b200_abi_version() == 1


def test_argument_validation_without_gpu():
    L = _lib.lib()
    a = (ctypes.c_float * 64)()
    p = ctypes.addressof(a)
    p16 = (p + 15) // 16 * 16
    # null pointer
    assert L.msda_forward_f32(None, p16, p16, p16, p16, 1, 4, 1, 4, 1, 1, 4, p16, None) == -1
    assert b"null" in L.msda_b200_last_error()
    # bad shape
    assert L.msda_forward_f32(p16, p16, p16, p16, p16, 1, 0, 1, 4, 1, 1, 4, p16, None) == -2
    assert L.msda_forward_f32(p16, p16, p16, p16, p16, 1, 4, 1, 4, 17, 1, 4, p16, None) == -2
    assert L.msda_forward_f32(p16, p16, p16, p16, p16, 1, 1 << 24, 8, 32, 1, 1, 4, p16, None) == -2
    # misaligned
    assert L.msda_forward_f32(p16 + 4, p16, p16, p16, p16, 1, 4, 1, 4, 1, 1, 4, p16, None) == -3
    assert L.msda_backward_f32(p16, p16, p16, p16, p16, p16, 1, 4, 1, 4, 1, 1, 4, None, p16, p16, 0, None) == -1
    # tuning knobs
    assert L.msda_b200_set_tuning(b"fwd_sample_batch", 3) == -4
    assert L.msda_b200_set_tuning(b"nope", 1) == -4
    old = L.msda_b200_get_tuning(b"fwd_passes")
    assert L.msda_b200_set_tuning(b"fwd_passes", 2) == 0 and L.msda_b200_get_tuning(b"fwd_passes") == 2
    L.msda_b200_set_tuning(b"fwd_passes", old)


def test_module_api_matches_reference_surface():
    with pytest.raises(ValueError):
        zb.MultiScaleDeformableAttention(embed_dim=30, num_heads=8)
    with pytest.warns(UserWarning):
        zb.MultiScaleDeformableAttention(embed_dim=24, num_heads=4)  # head dim 6: not a power of two
    m = zb.MultiScaleDeformableAttention()
    assert (m.embed_dim, m.num_heads, m.num_levels, m.num_points, m.im2col_step, m.batch_first) == (256, 8, 4, 4, 64, False)
    assert sorted(m.state_dict().keys()) == sorted(
        "%s.%s" % (a, b) for a in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
        for b in ("weight", "bias"))
    assert m.sampling_offsets.weight.shape == (256, 256) and m.attention_weights.weight.shape == (128, 256)
    # reference init (ms_deform_attn.py:194-217)
    assert m.sampling_offsets.weight.abs().max() == 0 and m.attention_weights.weight.abs().max() == 0
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1.0, 2.0, 3.0, 4.0])) and b[0, 0, :, 1].abs().max() < 1e-6
    assert torch.allclose(b[2, 1, 3], torch.tensor([0.0, 4.0]), atol=1e-5)
    m.freeze_sampling_offsets(); m.freeze_attention_weights()
    assert not m.sampling_offsets.weight.requires_grad and not m.attention_weights.bias.requires_grad
    m._reset_parameters()


def test_reference_state_dict_loads():
    g = load_golden("module_enc")
    C, M, L, P, bf = (int(x) for x in g["cfg"])
    m = zb.MultiScaleDeformableAttention(C, M, L, P, batch_first=bool(bf))
    sd = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    m.load_state_dict(sd, strict=True)


def test_no_cpu_fallback():
    m = zb.MultiScaleDeformableAttention(32, 2, 2, 2, batch_first=True)
    sh = torch.tensor([[3, 2], [1, 1]])
    with pytest.raises(RuntimeError, match="no CPU"):
        m(query=torch.randn(1, 7, 32), value=torch.randn(1, 7, 32), reference_points=torch.rand(1, 7, 2, 2),
          spatial_shapes=sh, level_start_index=torch.tensor([0, 6]))
    with pytest.raises(RuntimeError):
        zb.multi_scale_deformable_attn_pytorch()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        zb._C.ms_deform_attn_forward(torch.randn(1, 7, 2, 16), sh, torch.tensor([0, 6]),
                                     torch.rand(1, 7, 2, 2, 2, 2), torch.rand(1, 7, 2, 2, 2), 64)


def test_zira_rep_zero_linear_matches_reference_fixture():
    g = load_golden("zira_rep_linear")
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    m = zb.RepZeroLinear(24, 16).double()
    sd = {k[4:]: v for k, v in t.items() if k.startswith("pre.")}
    m.load_state_dict(sd)
    m.train()
    x = t["x"].clone().requires_grad_(True)
    out, loss = m(x)
    assert (out - t["train_out"]).abs().max() < 1e-13 and abs(float(loss.detach()) - float(t["train_loss"])) < 1e-13
    ((out * t["grad_out"]).sum() + loss * 0.1).backward()
    assert (x.grad - t["grad_x"]).abs().max() < 1e-12
    for k, p in m.named_parameters():
        assert (p.grad - t["pgrad." + k]).abs().max() < 1e-11, k
    m.eval()
    eo, el = m(t["x"])
    assert (eo - t["eval_out"]).abs().max() < 1e-13 and float(el) == 0
    m.__rep__()
    for k, v in m.state_dict().items():
        assert (v - t["post." + k]).abs().max() < 1e-15, k
    assert (m(t["x"])[0] - t["merged_eval_out"]).abs().max() < 1e-13
    m.train()
    mo, ml = m(t["x"])
    assert (mo - t["merged_train_out"]).abs().max() < 1e-13
    assert abs(float(ml) - float(t["merged_train_loss"])) < 1e-13
    # folded form == base linear + stand-alone adapter
    base = torch.nn.Linear(24, 16).double()
    y, l2 = m.forward_folded(t["x"], base.weight, base.bias)
    assert (y - (base(t["x"]) + mo)).abs().max() < 1e-13 and abs(float(l2) - float(ml)) < 1e-13
    m.eval()
    y, l2 = m.forward_folded(t["x"], base.weight, base.bias)
    assert l2 is None and (y - (base(t["x"]) + m(t["x"])[0])).abs().max() < 1e-13


@pytest.mark.parametrize("name", ["zira_rep_conv1x1", "zira_rep_conv3x3s2"])
def test_zira_rep_zero_conv2d_matches_reference_fixture(name):
    """Host logic of row N3 on CPU/fp64: the drop-in RepZeroConv2d and its rows (channels-last, im2col) fold."""
    g = load_golden(name)
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    cin, cout, groups, ksize, stride, padding = (int(v) for v in g["meta"])
    m = zb.RepZeroConv2d(cin, cout, kernel_size=ksize, stride=stride, padding=padding).double()
    m.load_state_dict({k[4:]: v for k, v in t.items() if k.startswith("pre.")})
    base = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, ksize, stride, padding), torch.nn.GroupNorm(groups, cout)).double()
    base.load_state_dict({k[5:]: v for k, v in t.items() if k.startswith("base.")})
    m.train()
    x = t["x"].clone().requires_grad_(True)
    out, loss = m(x)
    assert (out - t["train_out"]).abs().max() < 1e-13 and abs(float(loss.detach()) - float(t["train_loss"])) < 1e-13
    # the rows fold: GroupNorm(conv_0(x) + adapter(x)) on [N, HW, C] rows, reference lines :492-493
    from ziragroundingdino_b200.layer_ops import group_norm_rows
    assert m.rows_foldable(base[0])
    N, _, H, W = x.shape
    rows = x.flatten(2).transpose(1, 2)
    y, (ho, wo), loss2 = m.forward_folded_rows(rows, (H, W), base[0])
    src = group_norm_rows(y, base[1]).transpose(1, 2).reshape(N, cout, ho, wo)
    assert (src - t["train_src"]).abs().max() < 1e-11 and abs(float(loss2.detach()) - float(t["train_loss"])) < 1e-13
    ((src * t["grad_src"]).sum() + loss2 * 0.1).backward()
    assert (x.grad - t["grad_x"]).abs().max() < 1e-10
    for k, p in m.named_parameters():
        assert (p.grad - t["pgrad." + k]).abs().max() < 1e-10, k
    m.eval()
    eo, el = m(t["x"])
    assert (eo - t["eval_out"]).abs().max() < 1e-13 and float(el) == 0
    y, (ho, wo), l3 = m.forward_folded_rows(t["x"].flatten(2).transpose(1, 2), (H, W), base[0])
    src = group_norm_rows(y, base[1]).transpose(1, 2).reshape(N, cout, ho, wo)
    assert l3 is None and (src - t["eval_src"]).abs().max() < 1e-11
    m.__rep__()
    for k, v in m.state_dict().items():
        assert (v - t["post." + k]).abs().max() < 1e-15, k
    assert (m(t["x"])[0] - t["merged_eval_out"]).abs().max() < 1e-13
    m.train()
    mo, ml = m(t["x"])
    assert (mo - t["merged_train_out"]).abs().max() < 1e-13 and abs(float(ml) - float(t["merged_train_loss"])) < 1e-13


def test_zira_input_proj_keys_and_level_wiring():
    """ZiRaInputProj: the reference's parameter names (groundingdino_dual_zero_rep_branch.py:258-305) and level wiring
    (:483-523), checked against a literal transcription of that loop built from nn modules."""
    torch.manual_seed(0)
    m = zb.ZiRaInputProj(num_channels=(8, 12, 16), hidden_dim=32, num_feature_levels=5, norm_groups=4).double()
    keys = set(m.state_dict())
    for l in range(5):
        for k in ("input_proj.%d.0.weight", "input_proj.%d.0.bias", "input_proj.%d.1.weight", "input_proj.%d.1.bias",
                  "input_proj_conv_adapter.%d.weight", "input_proj_conv_adapter.%d.bias", "input_proj_conv_adapter.%d.scaling",
                  "input_proj_conv_adapter.%d.freeze_conv.weight", "input_proj_conv_adapter.%d.freeze_conv.bias"):
            assert k % l in keys
    assert len(keys) == 5 * 9
    assert m.input_proj[3][0].kernel_size == (3, 3) and m.input_proj[3][0].in_channels == 16
    assert m.input_proj[4][0].in_channels == 32 and m.input_proj_conv_adapter[4].stride == (2, 2)
    with torch.no_grad():
        for a in m.input_proj_conv_adapter:
            a.weight.normal_(0, 0.05); a.freeze_conv.weight.normal_(0, 0.05); a.freeze_conv.bias.normal_(0, 0.05)
    feats = [torch.randn(2, 8, 12, 10, dtype=torch.float64), torch.randn(2, 12, 6, 5, dtype=torch.float64),
             torch.randn(2, 16, 3, 3, dtype=torch.float64)]
    for training in (True, False):
        m.train(training)
        outs, loss = m(feats)
        want, wl = [], 0.0
        for l in range(5):
            x = feats[l] if l < 3 else (feats[-1] if l == 3 else want[-1])
            a, zl = m.input_proj_conv_adapter[l](x)
            want.append(m.input_proj[l][1](m.input_proj[l][0](x) + a))
            wl = wl + zl
        assert [tuple(o.shape) for o in outs] == [tuple(w.shape) for w in want]
        for o, w in zip(outs, want):
            assert (o - w).abs().max() < 1e-12
        assert abs(float(loss) - float(wl)) < 1e-12


def test_host_shapes_under_inference_mode():
    """inference tensors have no version counter; the shape cache must not touch ``_version`` for them."""
    import torch
    from ziragroundingdino_b200 import ms_deform_attn as M
    with torch.inference_mode():
        sh = torch.tensor([[4, 5], [2, 3]])
        assert sh.is_inference()
        assert M._host_shapes(sh) == ((4, 5), (2, 3))
        p = torch.nn.Parameter(torch.zeros(2))
    assert M._version_of(torch.zeros(2)) == 0
    with torch.inference_mode():
        t = torch.zeros(2)
        assert M._version_of(t) != M._version_of(t)       # never equal: no stale cache hit
    from ziragroundingdino_b200 import layer_ops
    with torch.inference_mode():
        w = torch.ones(3, 2)
        assert layer_ops.derived(w, "t").shape == (2, 3)


def test_fusedq_gate_needs_k_multiple_of_64():
    """3*M*L*P must be a multiple of the dgrad GEMM's K granularity for the fused-query backward."""
    from ziragroundingdino_b200 import fused
    assert fused.fusedq_ok(8, 4, 4, 32)
    assert not fused.fusedq_ok(6, 4, 4, 32)       # 288 columns: un-fused pair with the padded row
    assert not fused.fusedq_ok(8, 4, 4, 64)


def test_scaled_fp16_map_layout_and_scale_host_side():
    """Host logic of the opt-in scaled-fp16 accumulation (no GPU needed): the replicated map's row count for the Swin-T
    800x1333 levels, and the overflow bound behind f16acc_scale() -- for any gradient magnitude and query count the scale is
    a power of two with scale * max * Lq < 60000 (< the fp16 maximum 65504), and it is not needlessly small."""
    import ctypes
    import math
    import struct
    from ziragroundingdino_b200 import _lib
    L = _lib.lib()
    shapes = [(100, 167), (50, 84), (25, 42), (13, 21)]
    arr = (ctypes.c_int64 * 8)(*[d for hw in shapes for d in hw])
    assert L.msda_grad_value_h16_rows(arr, 4, 22223) == 16700 + 4200 + 4 * 1050 + 16 * 273       # replicas 1, 1, 4, 16
    assert L.msda_grad_value_h16_rows(arr, 4, 900) == 22223                                       # decoder: no replicas
    assert L.msda_grad_value_h16_rows(arr, 4, 0) == 0 and L.msda_grad_value_h16_rows(None, 4, 900) == 0
    tiny = (ctypes.c_int64 * 8)(6, 5, 3, 3, 2, 2, 1, 1)
    assert L.msda_grad_value_h16_rows(tiny, 4, 4000) == 32 * 30 + 64 * (9 + 4 + 1)               # 16*4000/30 = 2133 -> 32; the rest capped at 64
    bits = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]
    for lq in (1, 7, 900, 22223, 153520, 1 << 20):
        for e in range(-120, 121, 7):
            for m in (1.0, 1.37, 1.999):
                amax = m * 2.0 ** e
                s = L.msda_f16acc_scale(bits(amax), lq)
                assert s > 0 and math.frexp(s)[0] == 0.5, (amax, lq, s)                          # a power of two
                amax32 = struct.unpack("<f", struct.pack("<f", amax))[0]
                if amax32 * lq < 1e41:                  # beyond that (gradients near the fp32 maximum) the scale is clamped at 2^-126
                    assert s * amax32 * lq < 60000.0, (amax, lq, s)
                if 2.0 ** -100 < s < 2.0 ** 100:                                                   # not clamped: within 4x of the bound
                    assert s * amax32 * lq * 4.0 >= 60000.0 * 0.999, (amax, lq, s)
    for special in (0.0, float("inf"), float("nan")):
        assert L.msda_f16acc_scale(bits(special), 900) == 1.0
