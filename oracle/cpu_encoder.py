"""TEST INFRASTRUCTURE ONLY -- the reference's CPU path for the benchmark workload.

One image through ``layers`` deformable encoder layers, forward + backward, in fp32 on the host cores,
built from the oracle's restatement of the reference module (oracle/msda_oracle.py: module_forward with
the grid_sample core = ms_deform_attn.py:90-130, :281-350) plus the residual/LayerNorm/FFN of
transformer_for_adapter.py:876-907.  Used by bench.py for ``cpu_baseline`` and ``--impl reference``.
"""
import time

import torch
import torch.nn.functional as F

from . import msda_oracle as O


def make_layer_params(d_model=256, d_ffn=2048, M=8, L=4, P=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    return {
        "value_proj.weight": r(d_model, d_model, std=0.06), "value_proj.bias": torch.zeros(d_model),
        "sampling_offsets.weight": r(M * L * P * 2, d_model), "sampling_offsets.bias": r(M * L * P * 2, std=2.0),
        "attention_weights.weight": r(M * L * P, d_model), "attention_weights.bias": torch.zeros(M * L * P),
        "output_proj.weight": r(d_model, d_model, std=0.06), "output_proj.bias": torch.zeros(d_model),
        "linear1.weight": r(d_ffn, d_model, std=0.06), "linear1.bias": torch.zeros(d_ffn),
        "linear2.weight": r(d_model, d_ffn, std=0.02), "linear2.bias": torch.zeros(d_model),
        "norm1.weight": torch.ones(d_model), "norm1.bias": torch.zeros(d_model),
        "norm2.weight": torch.ones(d_model), "norm2.bias": torch.zeros(d_model),
    }


def encoder_layer(p, src, pos, refp, shapes, mask, M, L, P):
    src2 = O.module_forward(p, src + pos, src, mask, refp, shapes, M, L, P)
    src = F.layer_norm(src + src2, (src.shape[-1],), p["norm1.weight"], p["norm1.bias"])
    ffn = F.linear(F.relu(F.linear(src, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"], p["linear2.bias"])
    return F.layer_norm(src + ffn, (src.shape[-1],), p["norm2.weight"], p["norm2.bias"])


def timed_step(shapes, layers, threads, d_model=256, M=8, P=4, seed=0):
    """Returns seconds for one image through `layers` layers fwd+bwd on `threads` host threads."""
    torch.set_num_threads(threads)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(seed)
    sh = torch.tensor(shapes, dtype=torch.long)
    params = [make_layer_params(d_model, 2048, M, L, P, seed + i) for i in range(layers)]
    src = torch.randn(1, S, d_model, generator=g).requires_grad_(True)
    pos = torch.randn(1, S, d_model, generator=g)
    pts = []
    for (h, w) in shapes:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h) / h, torch.linspace(0.5, w - 0.5, w) / w, indexing="ij")
        pts.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    refp = torch.cat(pts, 0)[None, :, None, :].expand(1, S, L, 2).contiguous()
    t0 = time.perf_counter()
    x = src
    for p in params:
        x = encoder_layer(p, x, pos, refp, sh, None, M, L, P)
    x.square().mean().backward()
    return time.perf_counter() - t0


def timed_front(shapes, threads, channels=(192, 384, 768), d_model=256, seed=0):
    """Seconds for one image through the ZiRa-augmented input projection (oracle.input_proj_level per level:
    groundingdino_dual_zero_rep_branch.py:483-523) forward + backward w.r.t. the adapter parameters, fp32."""
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    nb = len(channels)
    feats = [torch.randn(1, c, h, w, generator=g) for c, (h, w) in zip(channels, shapes[:nb])]
    levels = []
    cin = channels[-1]
    for l in range(len(shapes)):
        c, k = (channels[l], 1) if l < nb else (cin, 3)
        base = (r(d_model, c, k, k), torch.zeros(d_model), torch.ones(d_model), torch.zeros(d_model))
        adapter = [t.requires_grad_(True) for t in (r(d_model, c, k, k), r(d_model), torch.full((1,), 0.1), r(d_model, c, k, k), r(d_model))]
        levels.append((base, adapter, k))
        if l >= nb:
            cin = d_model
    t0 = time.perf_counter()
    srcs, total = [], 0.0
    for l, (base, adapter, k) in enumerate(levels):
        x = feats[l] if l < nb else (feats[-1] if l == nb else srcs[-1])
        src, loss = O.input_proj_level(x, *base, 32, adapter, True, stride=1 if k == 1 else 2, padding=0 if k == 1 else 1)
        srcs.append(src)
        total = total + loss
    (sum(s.square().mean() for s in srcs) + 0.1 * total).backward()
    return time.perf_counter() - t0


def multihead_attention(p, prefix, query, key, value, n_heads, key_padding_mask=None, attn_mask=None):
    """nn.MultiheadAttention (seq-first, dropout 0) restated from its published definition -- PyTorch is third-party
    arithmetic on this path (the reference pins torch==2.4.0, requirements.txt:1): softmax(Q K^T / sqrt(d)) V with the
    packed in_proj and out_proj.  query [Lq, B, C], key/value [Lk, B, C]; key_padding_mask [B, Lk] bool (True = ignore)."""
    Lq, B, C = query.shape
    d = C // n_heads
    w, b = p[prefix + "in_proj_weight"], p[prefix + "in_proj_bias"]
    q = F.linear(query, w[:C], b[:C]).reshape(Lq, B * n_heads, d).transpose(0, 1)
    k = F.linear(key, w[C:2 * C], b[C:2 * C]).reshape(-1, B * n_heads, d).transpose(0, 1)
    v = F.linear(value, w[2 * C:], b[2 * C:]).reshape(-1, B * n_heads, d).transpose(0, 1)
    att = q @ k.transpose(1, 2) / (d ** 0.5)
    if attn_mask is not None:
        att = att.masked_fill(attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else att + attn_mask
    if key_padding_mask is not None:
        att = att.view(B, n_heads, Lq, -1).masked_fill(key_padding_mask[:, None, None, :], float("-inf")).view(B * n_heads, Lq, -1)
    out = (att.softmax(-1) @ v).transpose(0, 1).reshape(Lq, B, C)
    return F.linear(out, p[prefix + "out_proj.weight"], p[prefix + "out_proj.bias"])


def decoder_layer(p, tgt, query_pos, ref4, memory, memory_text, text_mask, shapes, mask, M, L, P):
    """transformer_for_adapter.py:1024-1073 with use_adapter=False, use_text_cross_attention=True, dropout 0.
    ``p`` uses the reference's state_dict keys.  tgt/query_pos [nq, B, C], ref4 [nq, B, L, 4], memory [S, B, C]."""
    C = tgt.shape[-1]
    ln = lambda x, n: F.layer_norm(x, (C,), p[n + ".weight"], p[n + ".bias"])
    q = tgt + query_pos
    tgt = ln(tgt + multihead_attention(p, "self_attn.", q, q, tgt, M), "norm2")
    if memory_text is not None:      # use_text_cross_attention
        text = memory_text.transpose(0, 1)
        tgt = ln(tgt + multihead_attention(p, "ca_text.", tgt + query_pos, text, text, M, key_padding_mask=text_mask), "catext_norm")
    attn_p = {k[len("cross_attn."):]: v for k, v in p.items() if k.startswith("cross_attn.")}
    tgt2 = O.module_forward(attn_p, (tgt + query_pos).transpose(0, 1), memory.transpose(0, 1), mask,
                            ref4.transpose(0, 1), shapes, M, L, P).transpose(0, 1)
    tgt = ln(tgt + tgt2, "norm1")
    ffn = F.linear(F.relu(F.linear(tgt, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"], p["linear2.bias"])
    return ln(tgt + ffn, "norm3")


def timed_decoder(shapes, layers, threads, nq=900, d_model=256, M=8, P=4, seed=0):
    """Seconds for one image's 900 queries through `layers` decoder layers (no text cross-attention) fwd+bwd, fp32;
    gradients flow to the queries and to the encoder memory as in the training step."""
    torch.set_num_threads(threads)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(seed)
    sh = torch.tensor(shapes, dtype=torch.long)
    r = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    params = []
    for i in range(layers):
        p = {("cross_attn." + k if k.split(".")[0] in ("value_proj", "sampling_offsets", "attention_weights", "output_proj") else k): v
             for k, v in make_layer_params(d_model, 2048, M, L, P, seed + i).items()}
        p.update({"self_attn.in_proj_weight": r(3 * d_model, d_model, std=0.06), "self_attn.in_proj_bias": torch.zeros(3 * d_model),
                  "self_attn.out_proj.weight": r(d_model, d_model, std=0.06), "self_attn.out_proj.bias": torch.zeros(d_model),
                  "norm3.weight": torch.ones(d_model), "norm3.bias": torch.zeros(d_model)})
        params.append(p)
    memory = torch.randn(S, 1, d_model, generator=g).requires_grad_(True)
    tgt = torch.randn(nq, 1, d_model, generator=g).requires_grad_(True)
    qpos = torch.randn(nq, 1, d_model, generator=g)
    ref4 = torch.cat([torch.rand(nq, 1, L, 2, generator=g) * 0.8 + 0.1, torch.rand(nq, 1, L, 2, generator=g) * 0.45 + 0.05], -1)
    t0 = time.perf_counter()
    x = tgt
    for p in params:
        x = decoder_layer(p, x, qpos, ref4, memory, None, None, sh, None, M, L, P)
    x.abs().mean().backward()
    return time.perf_counter() - t0


def bi_attention_block(p, v, l, mask_v, mask_l, num_heads):
    """The reference's BiAttentionBlock (fuse_modules.py:258-307 around BiMultiHeadAttention :154-254), dropout 0, restated
    with the attention matrix materialised exactly as the reference does: logits -> minus global max (:174) -> clamp
    +-50000 (:177-183) -> image<-text softmax over text (:207-213); transposed logits -> minus row max (:186-187) -> clamp
    (:188-195) -> mask -> text<-image softmax over image tokens (:198-205).  ``p`` uses the reference's state_dict keys."""
    C = v.shape[-1]
    v = F.layer_norm(v, (C,), p["layer_norm_v.weight"], p["layer_norm_v.bias"])
    l = F.layer_norm(l, (l.shape[-1],), p["layer_norm_l.weight"], p["layer_norm_l.bias"])
    lin = lambda x, n: F.linear(x, p["attn." + n + ".weight"], p["attn." + n + ".bias"])
    B, n_img, _ = v.shape
    E = p["attn.v_proj.weight"].shape[0]
    d = E // num_heads
    split = lambda t: t.view(B, -1, num_heads, d).transpose(1, 2).reshape(B * num_heads, -1, d)
    q, k = split(lin(v, "v_proj") * d ** -0.5), split(lin(l, "l_proj"))
    vv, vl = split(lin(v, "values_v_proj")), split(lin(l, "values_l_proj"))
    w = torch.bmm(q, k.transpose(1, 2))
    w = (w - w.max()).clamp(min=-50000, max=50000)
    wl = w.transpose(1, 2)
    wl = (wl - wl.max(dim=-1, keepdim=True)[0]).clamp(min=-50000, max=50000)
    if mask_v is not None:
        wl = wl.masked_fill(mask_v[:, None, None, :].repeat(1, num_heads, 1, 1).flatten(0, 1), float("-inf"))
    if mask_l is not None:
        w = w.masked_fill(mask_l[:, None, None, :].repeat(1, num_heads, 1, 1).flatten(0, 1), float("-inf"))
    out_v = torch.bmm(w.softmax(-1), vl).view(B, num_heads, n_img, d).transpose(1, 2).reshape(B, n_img, E)
    out_l = torch.bmm(wl.softmax(-1), vv).view(B, num_heads, -1, d).transpose(1, 2).reshape(B, -1, E)
    return v + p["gamma_v"] * lin(out_v, "out_v_proj"), l + p["gamma_l"] * lin(out_l, "out_l_proj")
