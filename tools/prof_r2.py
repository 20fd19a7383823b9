"""One launch of every round-2 kernel at config 2's shapes (4 images, Swin-T 800x1333, bf16 unless noted), for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'msda_bwd_vec|msda_scatter_mma|msda_fwd_vec|ffn_chain|linear_tf32x3|flatten_levels|encoder_proposals' \
        -o gpurun_out/r2_kernels python tools/prof_r2.py
Environment MSDA_B200_TUNING selects the scatter variant as usual."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ziragroundingdino_b200 as zb  # noqa: E402
from ziragroundingdino_b200 import blocks, fused, synthetic as syn, transformer_io as tio  # noqa: E402

dev = "cuda:0"
N = 4
shapes = syn.SWIN_T_800x1333
inp = syn.core_inputs(shapes, N, dtype=torch.bfloat16, regime="local", device=dev, seed=5)
a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
ref = syn.encoder_reference_points(shapes, torch.ones(N, 4, 2, device=dev), dev).contiguous()
for _ in range(2):
    o = zb._C.ms_deform_attn_forward(*a, 64)
    gv, dq = fused.backward_fusedq16(*a, inp["grad_out"], ref, 2)
R, C, F = N * sum(h * w for h, w in shapes), 256, 2048
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(R, C, device=dev, generator=g).bfloat16()
w1 = (torch.randn(F, C, device=dev, generator=g) * 0.06).bfloat16()
w2 = (torch.randn(C, F, device=dev, generator=g) * 0.02).bfloat16()
b1, b2 = torch.zeros(F, device=dev), torch.zeros(C, device=dev)
bits = torch.zeros((F // 32, R), dtype=torch.int32, device=dev)
for _ in range(2):
    y = blocks.ffn_chain_fwd16(x, w1, b1, w2, b2, bits)
    dx = blocks.ffn_chain_bwd16(x.clone(), w2.t().contiguous(), w1.t().contiguous(), bits)
x32 = torch.randn(R, C, device=dev, generator=g)
w32 = torch.randn(C, C, device=dev, generator=g) * 0.06
wq = torch.randn(384, C, device=dev, generator=g) * 0.05
ws, wqs = fused.split_tf32(w32), fused.split_tf32(wq)
bq = torch.zeros(384, device=dev)
for _ in range(2):
    y32 = fused.linear32(x32, ws, None)
    loc, aw = fused.query_proj32(x32, wqs, bq, ref.view(R, 4, 2), 2, inp["shapes"], 8, 4, 4)
srcs = [torch.randn(N, C, h, w, device=dev).bfloat16() for h, w in shapes]
poss = [torch.randn(N, C, h, w, device=dev).bfloat16() for h, w in shapes]
masks = [torch.zeros(N, h, w, dtype=torch.bool, device=dev) for h, w in shapes]
lvl = torch.randn(4, C, device=dev).bfloat16()
for _ in range(2):
    src, mask, pos, shp, sh, lsi, vr = tio.flatten_levels(srcs, masks, poss, lvl)
    om, op = tio.gen_encoder_output_proposals(src, mask, shapes)
torch.cuda.synchronize()
print("done")
