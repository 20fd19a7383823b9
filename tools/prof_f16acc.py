"""Two launches of the scaled-fp16 scatter (msda_backward_fusedq_h16) and of the two-operand GEMMs at config 2's shapes, for
    ncu --set full --clock-control none --import-source on -k regex:'msda_bwd_vec|linear_tc' -o gpurun_out/... python tools/prof_f16acc.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ziragroundingdino_b200 import blocks, fused, synthetic as syn  # noqa: E402

dev = "cuda:0"
N = 4
shapes = syn.SWIN_T_800x1333
inp = syn.core_inputs(shapes, N, dtype=torch.bfloat16, regime="local", device=dev, seed=5)
a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
ref = syn.encoder_reference_points(shapes, torch.ones(N, 4, 2, device=dev), dev).contiguous()
for _ in range(2):
    buf, dq = fused.backward_fusedq_h16(*a, inp["grad_out"], ref, 2)
R, C = N * sum(h * w for h, w in shapes), 256
g = torch.Generator(device=dev).manual_seed(1)
src = torch.randn(R, C, device=dev, generator=g).bfloat16()
pos = torch.randn(R, C, device=dev, generator=g).bfloat16()
wq = (torch.randn(384, C, device=dev, generator=g) * 0.05).bfloat16()
bq = torch.zeros(384, device=dev)
gv = torch.randn(R, C, device=dev, generator=g).bfloat16()
w12 = (torch.randn(C, C + 384, device=dev, generator=g) * 0.05).bfloat16()
acc = torch.randn(R, C, device=dev, generator=g).bfloat16()
for _ in range(2):
    loc, aw = fused.query_proj16(src, wq, bq, ref.view(R, 4, 2), 2, inp["shapes"], 8, 4, 4, q_add=pos)
    out = blocks.linear_accum2_16(gv, dq, w12, acc, torch.empty_like(acc))
wo = (torch.randn(C, C, device=dev, generator=g) * 0.06).bfloat16()
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(2):
    z, y, mean, rstd = blocks.linear_add_ln16(gv, wo, None, src, gam, bet, 1e-5)
torch.cuda.synchronize()
print("done")
