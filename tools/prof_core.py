"""Runs a few forward/backward launches of the core op for ncu: prof_core.py <f32|bf16> <N> <local|uniform> [Lq]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ziragroundingdino_b200 as zb  # noqa: E402
from ziragroundingdino_b200 import synthetic as syn  # noqa: E402

dt = {"f32": torch.float32, "bf16": torch.bfloat16}[sys.argv[1]]
N = int(sys.argv[2])
regime = sys.argv[3]
Lq = int(sys.argv[4]) if len(sys.argv) > 4 else None
inp = syn.core_inputs(syn.SWIN_T_800x1333, N, dtype=dt, regime=regime, Lq=Lq, device="cuda:0")
args = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
for _ in range(4):
    o = zb._C.ms_deform_attn_forward(*args, 64)
    g = zb._C.ms_deform_attn_backward(*args, inp["grad_out"], 64)
torch.cuda.synchronize()
print("done", o.shape)
