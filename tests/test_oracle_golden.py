"""CPU: pins the oracle (oracle/) against the fixtures the reference itself produced
(tests/golden/make_golden.py).  fp64 must agree to round-off; fp32 to the measured noise floor
between the reference's two formulations."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import msda_oracle as O

CORE = ["core_tiny_f64", "core_d32_f32", "core_l5_f32", "core_oddD_f32", "core_d64_f32"]


@pytest.mark.parametrize("name", CORE)
def test_c_oracle_f64_matches_reference(name):
    g = load_golden(name)
    v, l, a, go = (g[k].astype(np.float64) for k in ("value", "loc", "aw", "grad_out"))
    out = O.c_forward(v, g["shapes"], l, a)
    assert np.abs(out - g["out_f64"]).max() < 1e-12
    gv, gl, ga = O.c_backward(v, g["shapes"], l, a, go)
    assert rel_err(gv, g["grad_value_f64"]) < 1e-12
    assert rel_err(gl, g["grad_loc_f64"]) < 1e-11
    assert rel_err(ga, g["grad_aw_f64"]) < 1e-12


@pytest.mark.parametrize("name", [n for n in CORE if n.endswith("f32")])
def test_c_oracle_f32_matches_reference(name):
    g = load_golden(name)
    out = O.c_forward(g["value"], g["shapes"], g["loc"], g["aw"])
    assert out.dtype == np.float32
    # vs the reference's fp32 CPU path and vs fp64 truth: 1e-5 max-abs (north-star forward bar)
    assert np.abs(out - g["out"]).max() < 1e-5
    assert np.abs(out - g["out_f64"]).max() < 1e-5
    gv, gl, ga = O.c_backward(g["value"], g["shapes"], g["loc"], g["aw"], g["grad_out"])
    # 1e-4 relative (north-star backward bar)
    assert rel_err(gv, g["grad_value_f64"]) < 1e-4
    assert rel_err(gl, g["grad_loc_f64"]) < 1e-4
    assert rel_err(ga, g["grad_aw_f64"]) < 1e-4


@pytest.mark.parametrize("name", CORE)
def test_grid_sample_restatement_matches_reference(name):
    g = load_golden(name)
    v, l, a = (torch.from_numpy(g[k]).clone().requires_grad_(True) for k in ("value", "loc", "aw"))
    out = O.grid_sample_core(v, torch.from_numpy(g["shapes"]), l, a)
    assert torch.equal(out, torch.from_numpy(g["out"])) or np.abs(out.detach().numpy() - g["out"]).max() < 1e-6
    out.backward(torch.from_numpy(g["grad_out"]))
    assert rel_err(v.grad.numpy(), g["grad_value"]) < 1e-5
    assert rel_err(l.grad.numpy(), g["grad_loc"]) < 1e-5
    assert rel_err(a.grad.numpy(), g["grad_aw"]) < 1e-5


@pytest.mark.parametrize("name", ["module_enc", "module_dec", "module_seqfirst", "module_d32"])
@pytest.mark.parametrize("core", ["grid_sample", "c"])
def test_module_restatement_matches_reference(name, core):
    g = load_golden(name)
    C, M, L, P, _bf = (int(x) for x in g["cfg"])
    params = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}
    mask = torch.from_numpy(g["mask"]) if "mask" in g else None
    sh = torch.from_numpy(g["shapes"])

    def c_core(v, shapes, loc, aw):
        return torch.from_numpy(O.c_forward(v.numpy(), shapes.numpy(), loc.numpy(), aw.numpy()))

    out = O.module_forward(params, torch.from_numpy(g["query"]), torch.from_numpy(g["value"]), mask,
                           torch.from_numpy(g["reference_points"]), sh, M, L, P,
                           core=O.grid_sample_core if core == "grid_sample" else c_core)
    ref = g["out"] if g["cfg"][4] else np.swapaxes(g["out"], 0, 1)
    assert np.abs(out.numpy() - ref).max() < 1e-11


def test_rep_zero_linear_restatement_matches_reference():
    g = load_golden("zira_rep_linear")
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    pre = (t["pre.weight"], t["pre.bias"], t["pre.scaling"], t["pre.freeze_linear.weight"], t["pre.freeze_linear.bias"])
    out, loss = O.rep_zero_linear(t["x"], *pre, training=True)
    assert np.abs(out.numpy() - g["train_out"]).max() < 1e-13
    assert abs(float(loss) - float(g["train_loss"][0])) < 1e-13
    out_e, loss_e = O.rep_zero_linear(t["x"], *pre, training=False)
    assert np.abs(out_e.numpy() - g["eval_out"]).max() < 1e-13 and float(loss_e) == 0.0
    post = O.rep_merge(*pre)
    names = ["post.weight", "post.bias", "post.scaling", "post.freeze_linear.weight", "post.freeze_linear.bias"]
    for got, n in zip(post, names):
        assert np.abs(got.numpy() - g[n]).max() < 1e-15, n
    out_m, _ = O.rep_zero_linear(t["x"], *post, training=False)
    assert np.abs(out_m.numpy() - g["merged_eval_out"]).max() < 1e-13
    # merged eval == unmerged train up to the 1e-8 re-initialised branch
    assert np.abs(g["merged_eval_out"] - g["train_out"]).max() < 1e-12
    assert np.abs(g["merged_train_out"] - g["train_out"]).max() < 1e-6


@pytest.mark.parametrize("name", ["zira_rep_conv1x1", "zira_rep_conv3x3s2"])
def test_rep_zero_conv2d_restatement_matches_reference(name):
    g = load_golden(name)
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    cin, cout, groups, ksize, stride, padding = (int(v) for v in g["meta"])
    pre = (t["pre.weight"], t["pre.bias"], t["pre.scaling"], t["pre.freeze_conv.weight"], t["pre.freeze_conv.bias"])
    out, loss = O.rep_zero_conv2d(t["x"], *pre, training=True, stride=stride, padding=padding)
    assert np.abs(out.numpy() - g["train_out"]).max() < 1e-13
    assert abs(float(loss) - float(g["train_loss"][0])) < 1e-13
    out_e, loss_e = O.rep_zero_conv2d(t["x"], *pre, training=False, stride=stride, padding=padding)
    assert np.abs(out_e.numpy() - g["eval_out"]).max() < 1e-13 and float(loss_e) == 0.0
    base = (t["base.0.weight"], t["base.0.bias"], t["base.1.weight"], t["base.1.bias"], groups)
    src, _ = O.input_proj_level(t["x"], *base, pre, training=True, stride=stride, padding=padding)
    assert np.abs(src.numpy() - g["train_src"]).max() < 1e-12
    src_e, _ = O.input_proj_level(t["x"], *base, pre, training=False, stride=stride, padding=padding)
    assert np.abs(src_e.numpy() - g["eval_src"]).max() < 1e-12
    post = O.rep_merge(*pre, init_scale=0.1)      # vis_scale == lan_scale == 0.1 (:62-63)
    names = ["post.weight", "post.bias", "post.scaling", "post.freeze_conv.weight", "post.freeze_conv.bias"]
    for got, n in zip(post, names):
        assert np.abs(got.numpy() - g[n]).max() < 1e-15, n
    out_m, _ = O.rep_zero_conv2d(t["x"], *post, training=False, stride=stride, padding=padding)
    assert np.abs(out_m.numpy() - g["merged_eval_out"]).max() < 1e-13
    assert np.abs(g["merged_eval_out"] - g["train_out"]).max() < 1e-12


def _layer_params(g):
    return {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")}


def test_encoder_layer_restatement_matches_reference():
    """oracle.cpu_encoder.encoder_layer (the CPU arm of bench.py) against the reference's own
    DeformableTransformerEncoderLayer (fixture generated from transformer_for_adapter.py:809-907)."""
    from oracle import cpu_encoder
    g = load_golden("layer_encoder")
    C, FF, M, L, P = (int(v) for v in g["cfg"])
    p = _layer_params(g)
    p = {(k[len("self_attn."):] if k.startswith("self_attn.") else k): v for k, v in p.items()}
    t = lambda k: torch.from_numpy(g[k])
    out = cpu_encoder.encoder_layer(p, t("src"), t("pos"), t("reference_points"), t("shapes"), t("mask"), M, L, P)
    assert np.abs(out.numpy() - g["out"]).max() < 1e-11
    assert float(g["adapter_loss"][0]) == 0.0


def test_decoder_layer_restatement_matches_reference():
    from oracle import cpu_encoder
    g = load_golden("layer_decoder")
    C, FF, M, L, P = (int(v) for v in g["cfg"])
    t = lambda k: torch.from_numpy(g[k])
    out = cpu_encoder.decoder_layer(_layer_params(g), t("tgt"), t("query_pos"), t("reference_points"), t("memory"),
                                    t("memory_text"), t("text_mask"), t("shapes"), t("mask"), M, L, P)
    assert np.abs(out.numpy() - g["out"]).max() < 1e-11
