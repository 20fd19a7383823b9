"""Experiment: cost of the scatter/gather kernels as a function of the sampled level's size.
Queries keep the encoder's raster order over the real pyramid (Lq = 22 223); all four sampled levels
get the same shape, so the time per sample of a 13x21 level can be compared with a 100x167 one."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ziragroundingdino_b200 as zb  # noqa: E402
from ziragroundingdino_b200 import synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(iters):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def inputs(shape, N=4, M=8, D=32, P=4, L=4, dtype=torch.bfloat16):
    g = torch.Generator(device=dev).manual_seed(3)
    pyr = syn.SWIN_T_800x1333
    Lq = sum(h * w for h, w in pyr)
    ref = syn.encoder_reference_points(pyr, torch.ones(N, len(pyr), 2, device=dev), dev)[:, :, :1]   # [N,Lq,1,2]
    shapes = [shape] * L
    sh, lsi = syn.level_tensors(shapes, dev)
    S = L * shape[0] * shape[1]
    th = torch.arange(M, device=dev, dtype=torch.float32) * (2.0 * math.pi / M)
    d = torch.stack([th.cos(), th.sin()], -1)
    d = d / d.abs().max(-1, keepdim=True)[0]
    k = torch.arange(1, P + 1, device=dev, dtype=torch.float32)
    off = d[:, None, None, :] * k[None, None, :, None] + (torch.rand(N, Lq, M, L, P, 2, generator=g, device=dev) - 0.5)
    norm = torch.tensor([shape[1], shape[0]], device=dev, dtype=torch.float32)
    loc = (ref[:, :, None, :, None, :] + off / norm).contiguous()
    value = torch.randn(N, S, M, D, generator=g, device=dev).to(dtype)
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, device=dev), -1).view(N, Lq, M, L, P).contiguous()
    gout = torch.randn(N, Lq, M * D, generator=g, device=dev).to(dtype)
    return value, sh, lsi, loc, aw, gout


out = open(os.path.join(ROOT, "gpurun_out", "exp_levels.jsonl"), "w")
for dtype in (torch.bfloat16, torch.float32):
    for shape in syn.SWIN_T_800x1333 + [(200, 334)]:
        v, sh, lsi, loc, aw, go = inputs(shape, dtype=dtype)
        f = timeit(lambda: zb._C.ms_deform_attn_forward(v, sh, lsi, loc, aw, 64))
        b = timeit(lambda: zb._C.ms_deform_attn_backward(v, sh, lsi, loc, aw, go, 64))
        rec = dict(dtype=str(dtype), level=list(shape), fwd_us=f, bwd_us=b)
        print(rec); out.write(json.dumps(rec) + "\n")
