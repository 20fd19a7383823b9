"""Synthetic inputs of the reference's shapes (no datasets or checkpoints are reachable offline).

Level shapes follow SURVEY.md section 8: Swin-T 800x1333 -> 100x167, 50x84, 25x42, 13x21 (S = 22 223);
Swin-B 1024x1800 with 5 levels, strides 8..128 -> 128x225, 64x113, 32x57, 16x29, 8x15 (S = 38 440) or
strides 4..64 -> 256x450 ... 16x29 (S = 153 520).
"""
import math

import torch

SWIN_T_800x1333 = [(100, 167), (50, 84), (25, 42), (13, 21)]
SWIN_B_1024x1800_S8 = [(128, 225), (64, 113), (32, 57), (16, 29), (8, 15)]
SWIN_B_1024x1800_S4 = [(256, 450), (128, 225), (64, 113), (32, 57), (16, 29)]


def level_tensors(shapes, device):
    sh = torch.tensor(shapes, dtype=torch.long, device=device)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    return sh, lsi


def encoder_reference_points(shapes, valid_ratios, device):
    """Pixel-centre grid per level scaled by valid ratios -- what the reference's
    ``get_reference_points`` (transformer_for_adapter.py:483-497) produces.  valid_ratios [N, L, 2]
    (w, h).  Returns [N, S, L, 2]."""
    pts = []
    for lvl, (h, w) in enumerate(shapes):
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=device),
                                torch.linspace(0.5, w - 0.5, w, device=device), indexing="ij")
        ys = ys.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
        xs = xs.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
        pts.append(torch.stack((xs, ys), -1))
    ref = torch.cat(pts, 1)
    return ref[:, :, None] * valid_ratios[:, None]


def core_inputs(shapes, N, M=8, D=32, P=4, Lq=None, regime="local", dtype=torch.float32, device="cuda", seed=1):
    """Inputs of the core op. ``regime``: "local" = pixel-grid reference points plus the module's initial
    offset pattern (k-th point k pixels along the head's direction) with +-0.5 px jitter -- the
    pretrained-like case; "uniform" = locations ~ U(-0.05, 1.05), the worst case for cache locality.
    Lq defaults to S (encoder self-attention); otherwise queries get random reference points."""
    g = torch.Generator(device=device).manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    sh, lsi = level_tensors(shapes, device)
    value = torch.randn(N, S, M, D, generator=g, device=device).to(dtype)
    if Lq is None:
        Lq = S
    if regime == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g, device=device) * 1.1 - 0.05
    else:
        if Lq == S:
            ref = encoder_reference_points(shapes, torch.ones(N, L, 2, device=device), device)  # [N,S,L,2]
        else:
            ref = (torch.rand(N, Lq, 1, 2, generator=g, device=device) * 0.8 + 0.1).expand(N, Lq, L, 2)
        th = torch.arange(M, device=device, dtype=torch.float32) * (2.0 * math.pi / M)
        d = torch.stack([th.cos(), th.sin()], -1)
        d = d / d.abs().max(-1, keepdim=True)[0]                                   # [M,2]
        k = torch.arange(1, P + 1, device=device, dtype=torch.float32)
        off = d[:, None, None, :] * k[None, None, :, None]                          # [M,1,P,2] pixels
        off = off + (torch.rand(N, Lq, M, L, P, 2, generator=g, device=device) - 0.5)
        norm = torch.stack([sh[:, 1], sh[:, 0]], -1).float()                        # (W, H)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, device=device), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=g, device=device).to(dtype)
    return dict(value=value, shapes=sh, level_start=lsi, loc=loc.contiguous(), aw=aw.contiguous(), grad_out=gout,
                dims=(N, S, M, D, L, Lq, P))


def algorithmic_bytes(N, S, M, D, L, Lq, P, value_bytes):
    """G = gather bytes; fwd L2-algorithmic = G + Bl + Ba + Bo; HBM-compulsory =
    Bv + Bl + Ba + Bo; bwd L2 = 3G + 2Bl + 2Ba + Bo; bwd HBM-compulsory = 2Bv + 2Bl + 2Ba + Bo."""
    G = N * Lq * M * L * P * 4 * D * value_bytes
    Bv = N * S * M * D * value_bytes
    Bl = N * Lq * M * L * P * 2 * 4
    Ba = N * Lq * M * L * P * 4
    Bo = N * Lq * M * D * value_bytes
    return dict(G=G, fwd_l2=G + Bl + Ba + Bo, fwd_hbm=Bv + Bl + Ba + Bo, bwd_l2=3 * G + 2 * Bl + 2 * Ba + Bo,
                bwd_hbm=2 * Bv + 2 * Bl + 2 * Ba + Bo)
