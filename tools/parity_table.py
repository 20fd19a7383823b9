"""Parity table (SURVEY.md 8(c) protocol): for each configuration, max-abs distance of the forward output to fp64 truth for
the new op, the C oracle (fp32) and -- when oracle/_ref/ref_C.so is present -- the reference's own CUDA op, plus
new-vs-reference; and the relative error of the three gradients vs the same-precision C oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ziragroundingdino_b200 as zb  # noqa: E402
from oracle import build_ref, msda_oracle as O  # noqa: E402
from ziragroundingdino_b200 import synthetic as syn  # noqa: E402

dev = "cuda:0"
ref = build_ref.load()
CASES = [("config1 enc Swin-T N=1 local", syn.SWIN_T_800x1333, 1, None, "local"),
         ("config1 enc Swin-T N=1 uniform", syn.SWIN_T_800x1333, 1, None, "uniform"),
         ("config3 dec Swin-T N=4 Lq=900", syn.SWIN_T_800x1333, 4, 900, "uniform"),
         ("config5 enc Swin-B(s8) L=5 N=1", syn.SWIN_B_1024x1800_S8, 1, None, "local"),
         # the stride-4 reading SURVEY 8 lists first: S = 153 520, 157 MB bf16 value map at N=2 -- NOT L2-resident.  The C
         # oracle runs on image 0 only (N=1 slice of the same inputs) to keep the CPU side in seconds.
         ("config5 enc Swin-B(s4) L=5 N=2", syn.SWIN_B_1024x1800_S4, 2, None, "local")]
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
print("%-34s %-5s %11s %11s %11s %11s | %9s %9s %9s" % ("case", "dtype", "new-truth", "orcl-truth", "ref-truth", "new-ref", "gV rel", "gLoc rel", "gAw rel"))
for name, shapes, N, Lq, regime in CASES:
    for dt in (torch.float32, torch.bfloat16):
        inp = syn.core_inputs(shapes, N, dtype=dt, regime=regime, Lq=Lq, device=dev, seed=1)
        args = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
        out = zb._C.ms_deform_attn_forward(*args, 64)
        gv, gl, ga = zb._C.ms_deform_attn_backward(*args, inp["grad_out"], 64)
        if "(s4)" in name:   # oracle on image 0: the op is independent per image
            out, gv, gl, ga = out[:1], gv[:1], gl[:1], ga[:1]
            inp = {k: (v[:1] if k in ("value", "loc", "aw", "grad_out") else v) for k, v in inp.items()}
            args = (inp["value"].contiguous(), inp["shapes"], inp["level_start"], inp["loc"].contiguous(), inp["aw"].contiguous())
        c = {k: inp[k].cpu() for k in ("value", "shapes", "loc", "aw", "grad_out")}
        v64 = c["value"].double().numpy()
        truth = O.c_forward(v64, c["shapes"].numpy(), c["loc"].double().numpy(), c["aw"].double().numpy())
        o32 = O.c_forward(c["value"].float().numpy(), c["shapes"].numpy(), c["loc"].numpy(), c["aw"].numpy())
        ogv, ogl, oga = O.c_backward(c["value"].float().numpy(), c["shapes"].numpy(), c["loc"].numpy(), c["aw"].numpy(),
                                     c["grad_out"].float().numpy())
        d_new = np.abs(out.float().cpu().numpy() - truth).max()
        d_or = np.abs(o32 - truth).max()
        d_ref = d_nr = float("nan")
        if ref is not None and dt == torch.float32:
            r_out = ref.ms_deform_attn_forward(*args, 64)
            d_ref = np.abs(r_out.cpu().numpy() - truth).max()
            d_nr = (r_out - out).abs().max().item()
        bad = (np.abs(gl.cpu().numpy() - ogl) > 1e-4 * np.abs(ogl).max()).mean()
        print("%-34s %-5s %11.3e %11.3e %11.3e %11.3e | %9.2e %9.2e %9.2e  (gLoc slots off: %.1e)" % (
            name, "f32" if dt == torch.float32 else "bf16", d_new, d_or, d_ref, d_nr, rel(gv.float().cpu().numpy(), ogv),
            rel(gl.cpu().numpy(), ogl), rel(ga.cpu().numpy(), oga), bad))
