"""Generates tests/golden/*.npz by running the REFERENCE's own Python code in the build container.

Run once here (``python tests/golden/make_golden.py``); the fixtures are committed because
/root/reference does not exist on the GPU box.  Nothing is copied from the reference: its hot-path
file is imported by path with a stub ``groundingdino._C`` (the package import chain needs
detectron2), and the ZiRa ``RepZeroLinear`` class is executed from its AST because the file that
defines it imports detectron2 at module scope.

Sources exercised (relative to /root/reference/groundingdino/models/GroundingDINO/):
  ms_deform_attn.py:90-130   multi_scale_deformable_attn_pytorch  (+ torch autograd for the grads)
  ms_deform_attn.py:133-355  MultiScaleDeformableAttention (CPU branch, :345-348)
  groundingdino_dual_zero_rep_branch.py:62-64, :105-135  RepZeroLinear (train / eval / __rep__)
  groundingdino_dual_zero_rep_branch.py:64-103, :492-493 RepZeroConv2d and its use beside input_proj
  transformer_for_adapter.py:809-907, :910-1073  DeformableTransformerEncoderLayer / DecoderLayer (use_adapter=False)
  utils.py:56-116, transformer_for_adapter.py:216-262, :311-329  proposals, valid ratios, flatten, top-k selection
  fuse_modules.py:99-307  BiMultiHeadAttention / BiAttentionBlock (image <-> text fusion)
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_DIR = "/root/reference/groundingdino/models/GroundingDINO"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_msda():
    pkg = types.ModuleType("groundingdino")
    pkg.__path__ = []
    pkg._C = types.ModuleType("groundingdino._C")
    sys.modules["groundingdino"] = pkg
    sys.modules["groundingdino._C"] = pkg._C
    spec = importlib.util.spec_from_file_location("ref_msda", os.path.join(REF_DIR, "ms_deform_attn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_rep_zero_linear(name="RepZeroLinear"):
    path = os.path.join(REF_DIR, "groundingdino_dual_zero_rep_branch.py")
    tree = ast.parse(open(path).read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            keep.append(node)
        if isinstance(node, ast.Assign) and any(
            isinstance(t, ast.Name) and t.id in ("zero_value", "lan_scale", "vis_scale") for t in node.targets
        ):
            keep.append(node)
    from torch.nn.common_types import _size_2_t
    ns = {"torch": torch, "nn": torch.nn, "Tensor": torch.Tensor, "_size_2_t": _size_2_t}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns[name]


def core_case(ref, seed, N, shapes, M, D, Lq, P, dtype, lo=-0.1, hi=1.1, special=False):
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * (hi - lo) + lo
    if special:
        # exact pixel centres, exact borders and far-outside points (guards at im2col.cuh:288, :56-78)
        flat = loc.view(-1, 2)
        flat[0] = torch.tensor([0.0, 0.0])
        flat[1] = torch.tensor([1.0, 1.0])
        flat[2] = torch.tensor([0.5, 0.5])
        flat[3] = torch.tensor([-3.0, 0.5])
        flat[4] = torch.tensor([0.5, 7.0])
        flat[5] = torch.tensor([0.5 / shapes[0][1], 0.5 / shapes[0][0]])  # centre of pixel (0,0), level 0
        flat[6] = torch.tensor([1.0 - 1e-9, 1e-9])
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, dtype=torch.float64), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)
    value, loc, aw, gout = (t.to(dtype) for t in (value, loc, aw, gout))
    sh = torch.tensor(shapes, dtype=torch.long)
    res = {}
    for tag, dt in (("", dtype), ("_f64", torch.float64)):
        v, l_, a = (t.to(dt).clone().requires_grad_(True) for t in (value, loc, aw))
        out = ref.multi_scale_deformable_attn_pytorch(v, sh, l_, a)
        out.backward(gout.to(dt))
        res.update({"out" + tag: out.detach().numpy(), "grad_value" + tag: v.grad.numpy(),
                    "grad_loc" + tag: l_.grad.numpy(), "grad_aw" + tag: a.grad.numpy()})
    res.update(value=value.numpy(), loc=loc.numpy(), aw=aw.numpy(), grad_out=gout.numpy(),
               shapes=np.asarray(shapes, dtype=np.int64))
    return res


def module_case(ref, seed, C, M, L, P, shapes, N, Lq, ref_dim, batch_first, with_mask, self_attn):
    torch.manual_seed(seed)
    mod = ref.MultiScaleDeformableAttention(C, M, L, P, batch_first=batch_first).double()
    with torch.no_grad():  # make offsets / weights query-dependent (init has zero weights)
        mod.sampling_offsets.weight.normal_(0, 0.05)
        mod.attention_weights.weight.normal_(0, 0.3)
        mod.attention_weights.bias.normal_(0, 0.3)
        mod.value_proj.bias.normal_(0, 0.1)
        mod.output_proj.bias.normal_(0, 0.1)
    S = sum(h * w for h, w in shapes)
    if self_attn:
        Lq = S
    value = torch.randn(N, S, C, dtype=torch.float64)
    query = torch.randn(N, Lq, C, dtype=torch.float64)
    if ref_dim == 2:
        refp = torch.rand(N, Lq, L, 2, dtype=torch.float64)
    else:
        refp = torch.cat([torch.rand(N, Lq, L, 2, dtype=torch.float64) * 0.8 + 0.1,
                          torch.rand(N, Lq, L, 2, dtype=torch.float64) * 0.45 + 0.05], -1)
    mask = None
    if with_mask:
        mask = torch.zeros(N, S, dtype=torch.bool)
        mask[0, S // 2:] = True
        mask[-1, ::7] = True
    sh = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    q_in = query.clone().requires_grad_(True)
    v_in = value.clone().requires_grad_(True)
    qa, va = (q_in, v_in) if batch_first else (q_in.transpose(0, 1), v_in.transpose(0, 1))
    out = mod(query=qa, value=va, key_padding_mask=mask, reference_points=refp, spatial_shapes=sh,
              level_start_index=lsi)
    gout = torch.randn_like(out)
    out.backward(gout)
    res = {"query": query.numpy(), "value": value.numpy(), "reference_points": refp.numpy(),
           "shapes": sh.numpy(), "out": out.detach().numpy(), "grad_out": gout.numpy(),
           "grad_query": q_in.grad.numpy(), "grad_value_in": v_in.grad.numpy(),
           "cfg": np.asarray([C, M, L, P, int(batch_first)], dtype=np.int64)}
    if mask is not None:
        res["mask"] = mask.numpy()
    for k, v in mod.state_dict().items():
        res["param." + k] = v.numpy()
    for k, p in mod.named_parameters():
        res["pgrad." + k] = p.grad.numpy()
    return res


def zira_case(RepZeroLinear, seed):
    torch.manual_seed(seed)
    m = RepZeroLinear(24, 16).double()
    with torch.no_grad():  # non-trivial branch and accumulated soft-frozen weights
        m.weight.normal_(0, 0.05)
        m.bias.normal_(0, 0.05)
        m.freeze_linear.weight.normal_(0, 0.2)
        m.freeze_linear.bias.normal_(0, 0.2)
        m.scaling.fill_(0.3)
    x = torch.randn(5, 7, 24, dtype=torch.float64)
    res = {"x": x.numpy()}
    for k, v in m.state_dict().items():
        res["pre." + k] = v.numpy().copy()
    m.train()
    xi = x.clone().requires_grad_(True)
    out, loss = m(xi)
    gout = torch.randn_like(out)
    (out * gout).sum().add(loss * 0.1).backward()
    res.update(train_out=out.detach().numpy(), train_loss=loss.detach().numpy().reshape(1),
               grad_out=gout.numpy(), grad_x=xi.grad.numpy())
    for k, p in m.named_parameters():
        res["pgrad." + k] = p.grad.numpy().copy()
    m.eval()
    eo, el = m(x)
    res.update(eval_out=eo.detach().numpy(), eval_loss=el.detach().numpy().reshape(1))
    m.__rep__()
    for k, v in m.state_dict().items():
        res["post." + k] = v.detach().numpy().copy()
    m.eval()
    res["merged_eval_out"] = m(x)[0].detach().numpy()
    m.train()
    mo, ml = m(x)
    res.update(merged_train_out=mo.detach().numpy(), merged_train_loss=ml.detach().numpy().reshape(1))
    return res


def zira_conv_case(RepZeroConv2d, seed, cin, cout, groups, ksize, stride, padding, hw):
    """RepZeroConv2d (groundingdino_dual_zero_rep_branch.py:64-103) alone, and inside the input projection of one
    level exactly as the reference composes it (:492-493): GroupNorm(conv_0(x) + adapter(x))."""
    torch.manual_seed(seed)
    m = RepZeroConv2d(cin, cout, kernel_size=ksize, stride=stride, padding=padding).double()
    base = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, kernel_size=ksize, stride=stride, padding=padding),
                               torch.nn.GroupNorm(groups, cout)).double()
    with torch.no_grad():
        m.weight.normal_(0, 0.05)
        m.bias.normal_(0, 0.05)
        m.freeze_conv.weight.normal_(0, 0.1)
        m.freeze_conv.bias.normal_(0, 0.1)
        m.scaling.fill_(0.3)
        base[1].weight.normal_(1, 0.2)
        base[1].bias.normal_(0, 0.2)
    x = torch.randn(2, cin, *hw, dtype=torch.float64)
    res = {"x": x.numpy(), "meta": np.array([cin, cout, groups, ksize, stride, padding])}
    for k, v in m.state_dict().items():
        res["pre." + k] = v.numpy().copy()
    for k, v in base.state_dict().items():
        res["base." + k] = v.numpy().copy()
    m.train()
    xi = x.clone().requires_grad_(True)
    out, loss = m(xi)
    src = base[1](base[0](xi) + out)                       # :493
    gsrc = torch.randn_like(src)
    (src * gsrc).sum().add(loss * 0.1).backward()
    res.update(train_out=out.detach().numpy(), train_loss=loss.detach().numpy().reshape(1), train_src=src.detach().numpy(),
               grad_src=gsrc.numpy(), grad_x=xi.grad.numpy())
    for k, p in m.named_parameters():
        res["pgrad." + k] = p.grad.numpy().copy()
    m.eval()
    eo, el = m(x)
    res.update(eval_out=eo.detach().numpy(), eval_loss=el.detach().numpy().reshape(1),
               eval_src=base[1](base[0](x) + eo).detach().numpy())
    m.__rep__()
    for k, v in m.state_dict().items():
        res["post." + k] = v.detach().numpy().copy()
    m.eval()
    res["merged_eval_out"] = m(x)[0].detach().numpy()
    m.train()
    mo, ml = m(x)
    res.update(merged_train_out=mo.detach().numpy(), merged_train_loss=ml.detach().numpy().reshape(1))
    return res


def load_reference_layers(ref):
    """DeformableTransformerEncoderLayer / DecoderLayer (transformer_for_adapter.py:809-907, :910-1073) executed from
    their AST (the file imports detectron2-dependent modules at the top), with MSDeformAttn bound to the reference's own
    MultiScaleDeformableAttention (CPU branch) and _get_activation_fn taken from the reference's utils.py:188-201."""
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "Tensor": torch.Tensor,
          "Optional": __import__("typing").Optional, "MSDeformAttn": ref.MultiScaleDeformableAttention}
    tree = ast.parse(open(os.path.join(REF_DIR, "utils.py")).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_get_activation_fn"]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "utils.py", "exec"), ns)
    tree = ast.parse(open(os.path.join(REF_DIR, "transformer_for_adapter.py")).read())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef)
            and n.name in ("DeformableTransformerEncoderLayer", "DeformableTransformerDecoderLayer")]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "transformer_for_adapter.py", "exec"), ns)
    return ns["DeformableTransformerEncoderLayer"], ns["DeformableTransformerDecoderLayer"]


def _randomise(layer, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in layer.named_parameters():
            if "sampling_offsets.bias" in n:
                continue                                   # keep the directional grid init
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g, dtype=p.dtype) * (0.05 if "sampling_offsets" in n else 0.2))
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g, dtype=p.dtype))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g, dtype=p.dtype))


def layer_cases(ref, seed):
    Enc, Dec = load_reference_layers(ref)
    C, FF, M, L, P = 32, 64, 4, 3, 2
    shapes = [(5, 6), (3, 3), (2, 2)]
    S = sum(h * w for h, w in shapes)
    N, nq, ntext = 2, 7, 5
    torch.manual_seed(seed)
    sh = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    mask = torch.zeros(N, S, dtype=torch.bool)
    mask[1, S - 9:] = True
    out = {}
    # ---- encoder layer (use_adapter=False as in GroundingDINO_SwinT_OGC_rep.py:56) ----
    enc = Enc(C, FF, 0.0, "relu", L, M, P, use_adapter=False).double()
    _randomise(enc, seed + 1)
    src = torch.randn(N, S, C, dtype=torch.float64, requires_grad=True)
    pos = torch.randn(N, S, C, dtype=torch.float64)
    refp = torch.rand(N, S, L, 2, dtype=torch.float64)
    y, aloss = enc(src, pos, refp, sh, lsi, mask)
    gy = torch.randn_like(y)
    y.backward(gy)
    res = {"src": src.detach().numpy(), "pos": pos.numpy(), "reference_points": refp.numpy(), "shapes": sh.numpy(),
           "mask": mask.numpy(), "out": y.detach().numpy(), "adapter_loss": aloss.detach().numpy(), "grad_out": gy.numpy(),
           "grad_src": src.grad.numpy(), "cfg": np.asarray([C, FF, M, L, P], dtype=np.int64)}
    for k, v in enc.state_dict().items():
        res["param." + k] = v.numpy()
    out["layer_encoder"] = res
    # ---- decoder layer (text cross-attention on, as in the GroundingDINO config; seq-first tensors) ----
    dec = Dec(C, FF, 0.0, "relu", L, M, P, use_text_cross_attention=True, use_adapter=False).double()
    _randomise(dec, seed + 2)
    tgt = torch.randn(nq, N, C, dtype=torch.float64, requires_grad=True)
    qpos = torch.randn(nq, N, C, dtype=torch.float64)
    ref4 = torch.cat([torch.rand(nq, N, L, 2, dtype=torch.float64) * 0.8 + 0.1,
                      torch.rand(nq, N, L, 2, dtype=torch.float64) * 0.45 + 0.05], -1)
    memory = torch.randn(S, N, C, dtype=torch.float64, requires_grad=True)
    text = torch.randn(N, ntext, C, dtype=torch.float64)
    tmask = torch.zeros(N, ntext, dtype=torch.bool)
    tmask[0, -2:] = True
    y, aloss = dec(tgt=tgt, tgt_query_pos=qpos, tgt_reference_points=ref4, memory_text=text, text_attention_mask=tmask,
                   memory=memory, memory_key_padding_mask=mask, memory_level_start_index=lsi, memory_spatial_shapes=sh)
    gy = torch.randn_like(y)
    y.backward(gy)
    res = {"tgt": tgt.detach().numpy(), "query_pos": qpos.numpy(), "reference_points": ref4.numpy(), "memory": memory.detach().numpy(),
           "memory_text": text.numpy(), "text_mask": tmask.numpy(), "shapes": sh.numpy(), "mask": mask.numpy(),
           "out": y.detach().numpy(), "adapter_loss": aloss.detach().numpy(), "grad_out": gy.numpy(),
           "grad_tgt": tgt.grad.numpy(), "grad_memory": memory.grad.numpy(), "cfg": np.asarray([C, FF, M, L, P], dtype=np.int64)}
    for k, v in dec.state_dict().items():
        res["param." + k] = v.numpy()
    out["layer_decoder"] = res
    return out


def transformer_io_case(seed):
    """gen_encoder_output_proposals (utils.py:56-116) and Transformer.get_valid_ratio (transformer_for_adapter.py:216-223)
    executed from their AST, plus the inline flatten / top-k code of Transformer.forward (:238-262, :311-329) run on the
    same inputs as literal statements (they are not separate functions in the reference)."""
    ns = {"torch": torch, "Tensor": torch.Tensor}
    tree = ast.parse(open(os.path.join(REF_DIR, "utils.py")).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "gen_encoder_output_proposals"]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "utils.py", "exec"), ns)
    tree = ast.parse(open(os.path.join(REF_DIR, "transformer_for_adapter.py")).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "Transformer":
            keep = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name == "get_valid_ratio"]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "transformer_for_adapter.py", "exec"), ns)
    torch.manual_seed(seed)
    N, C, nq = 2, 8, 5
    shapes = [(6, 8), (3, 4), (2, 2)]
    masks = []
    for (h, w), (vh, vw) in zip(shapes, [(6, 8), (3, 4), (2, 2)]):
        m = torch.zeros(N, h, w, dtype=torch.bool)
        m[1, max(1, int(h * 0.7)):, :] = True
        m[1, :, max(1, int(w * 0.6)):] = True
        masks.append(m)
    srcs = [torch.randn(N, C, h, w, dtype=torch.float64) for h, w in shapes]
    poss = [torch.randn(N, C, h, w, dtype=torch.float64) for h, w in shapes]
    level_embed = torch.randn(len(shapes), C, dtype=torch.float64)
    # :238-262
    src_flatten, mask_flatten, lvl_pos = [], [], []
    for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, poss)):
        src_flatten.append(src.flatten(2).transpose(1, 2))
        mask_flatten.append(mask.flatten(1))
        lvl_pos.append(pos_embed.flatten(2).transpose(1, 2) + level_embed[lvl].view(1, 1, -1))
    src_flatten, mask_flatten, lvl_pos = torch.cat(src_flatten, 1), torch.cat(mask_flatten, 1), torch.cat(lvl_pos, 1)
    spatial_shapes = torch.as_tensor(shapes, dtype=torch.long)
    level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
    valid_ratios = torch.stack([ns["get_valid_ratio"](None, m) for m in masks], 1)
    res = {"shapes": spatial_shapes.numpy(), "level_embed": level_embed.numpy(), "src_flatten": src_flatten.numpy(),
           "mask_flatten": mask_flatten.numpy(), "lvl_pos_embed_flatten": lvl_pos.numpy(),
           "level_start_index": level_start_index.numpy(), "valid_ratios": valid_ratios.numpy()}
    for i, (s_, m_, p_) in enumerate(zip(srcs, masks, poss)):
        res["src%d" % i], res["mask%d" % i], res["pos%d" % i] = s_.numpy(), m_.numpy(), p_.numpy()
    memory = torch.randn(N, src_flatten.shape[1], C, dtype=torch.float64)
    for name, lw in (("", None), ("_learnedwh", torch.tensor([0.3, -0.2], dtype=torch.float64))):
        om, op = ns["gen_encoder_output_proposals"](memory, mask_flatten, spatial_shapes, lw)
        res["output_memory" + name], res["output_proposals" + name] = om.numpy(), op.numpy()
    res["memory"], res["learnedwh"] = memory.numpy(), np.array([0.3, -0.2])
    # :311-329
    om, op = ns["gen_encoder_output_proposals"](memory, mask_flatten, spatial_shapes)
    logits = torch.randn(N, src_flatten.shape[1], 6, dtype=torch.float64)
    coord = torch.randn(N, src_flatten.shape[1], 4, dtype=torch.float64) + op.masked_fill(op.isinf(), 0.0)
    topk_logits = logits.max(-1)[0]
    topk_proposals = torch.topk(topk_logits, nq, dim=1)[1]
    res.update(class_logits=logits.numpy(), coord_unselected=coord.numpy(), topk_proposals=topk_proposals.numpy(),
               refpoint_embed_undetach=torch.gather(coord, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4)).numpy(),
               init_box_proposal=torch.gather(op, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4)).sigmoid().numpy(),
               tgt_undetach=torch.gather(om, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, C)).numpy())
    return res


def bi_attention_case(seed):
    """BiAttentionBlock / BiMultiHeadAttention (fuse_modules.py:99-307) executed from their AST (the file imports timm,
    which is absent; DropPath is only instantiated for drop_path > 0)."""
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "DropPath": None}
    tree = ast.parse(open(os.path.join(REF_DIR, "fuse_modules.py")).read())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("BiMultiHeadAttention", "BiAttentionBlock")]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "fuse_modules.py", "exec"), ns)
    torch.manual_seed(seed)
    C, E, H, B, n_img, n_text = 32, 64, 4, 2, 41, 7
    blk = ns["BiAttentionBlock"](v_dim=C, l_dim=C, embed_dim=E, num_heads=H, dropout=0.0, drop_path=0.0).double()
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.startswith("gamma"):
                p.copy_(0.5 + 0.1 * torch.randn_like(p))          # layer scale large enough to matter
            elif n.endswith("bias"):
                p.copy_(0.1 * torch.randn_like(p))
            elif "layer_norm" in n:
                p.copy_(1 + 0.1 * torch.randn_like(p))
    v = torch.randn(B, n_img, C, dtype=torch.float64, requires_grad=True)
    l = torch.randn(B, n_text, C, dtype=torch.float64, requires_grad=True)
    mask_v = torch.zeros(B, n_img, dtype=torch.bool)
    mask_v[1, -9:] = True
    mask_l = torch.zeros(B, n_text, dtype=torch.bool)
    mask_l[0, -2:] = True
    ov, ol = blk(v, l, attention_mask_v=mask_v.clone(), attention_mask_l=mask_l.clone())
    gv, gl = torch.randn_like(ov), torch.randn_like(ol)
    ((ov * gv).sum() + (ol * gl).sum()).backward()
    res = {"v": v.detach().numpy(), "l": l.detach().numpy(), "mask_v": mask_v.numpy(), "mask_l": mask_l.numpy(),
           "out_v": ov.detach().numpy(), "out_l": ol.detach().numpy(), "grad_out_v": gv.numpy(), "grad_out_l": gl.numpy(),
           "grad_v": v.grad.numpy(), "grad_l": l.grad.numpy(), "cfg": np.asarray([C, E, H], dtype=np.int64)}
    for k, t in blk.state_dict().items():
        res["param." + k] = t.numpy()
    ov2, ol2 = blk(v.detach(), l.detach())                       # no masks
    res["out_v_nomask"], res["out_l_nomask"] = ov2.detach().numpy(), ol2.detach().numpy()
    return res


def main():
    ref = load_reference_msda()
    cases = {
        "core_tiny_f64": core_case(ref, 11, 2, [(6, 5), (3, 3)], 2, 4, 7, 2, torch.float64, special=True),
        "core_d32_f32": core_case(ref, 12, 1, [(9, 11), (5, 6), (3, 3), (2, 2)], 2, 32, 24, 4, torch.float32,
                                  special=True),
        "core_l5_f32": core_case(ref, 13, 2, [(6, 9), (3, 5), (2, 3), (1, 2), (1, 1)], 2, 32, 9, 4, torch.float32,
                                 lo=-0.3, hi=1.3),
        "core_oddD_f32": core_case(ref, 14, 1, [(5, 7), (3, 4)], 3, 6, 13, 3, torch.float32),
        "core_d64_f32": core_case(ref, 15, 1, [(5, 6), (3, 3)], 2, 64, 7, 4, torch.float32),
        "module_enc": module_case(ref, 21, 32, 4, 3, 2, [(5, 6), (3, 3), (2, 2)], 2, 0, 2, True, True, True),
        "module_dec": module_case(ref, 22, 32, 4, 3, 2, [(5, 6), (3, 3), (2, 2)], 2, 10, 4, True, True, False),
        "module_seqfirst": module_case(ref, 23, 32, 2, 2, 4, [(5, 4), (3, 2)], 3, 6, 2, False, False, False),
        "module_d32": module_case(ref, 24, 64, 2, 4, 4, [(6, 8), (3, 4), (2, 2), (1, 1)], 1, 0, 2, True, False, True),
    }
    cases["zira_rep_linear"] = zira_case(load_reference_rep_zero_linear(), 31)
    conv = load_reference_rep_zero_linear("RepZeroConv2d")
    cases["zira_rep_conv1x1"] = zira_conv_case(conv, 32, 24, 16, 4, 1, 1, 0, (5, 6))
    cases["zira_rep_conv3x3s2"] = zira_conv_case(conv, 33, 12, 16, 4, 3, 2, 1, (7, 6))
    cases.update(layer_cases(ref, 41))
    cases["transformer_io"] = transformer_io_case(51)
    cases["bi_attention"] = bi_attention_case(61)
    total = 0
    for name, arrs in cases.items():
        p = os.path.join(OUT, name + ".npz")
        np.savez_compressed(p, **arrs)
        total += os.path.getsize(p)
        print(name, os.path.getsize(p))
    print("total bytes", total)


if __name__ == "__main__":
    main()
