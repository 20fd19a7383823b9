"""bwd_passes sweep (consecutive unit tiles per CTA) for the fp32-accumulating and the scaled-fp16 scatter at config 2's launch."""
import json
import sys

import torch

sys.path.insert(0, ".")
from ziragroundingdino_b200 import _lib, fused, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
KN = 4


def kernel_us(fns, reps=12):
    for f in fns:
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(reps):
        fns[i % len(fns)]()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


sets = [syn.core_inputs(syn.SWIN_T_800x1333, KN, dtype=torch.bfloat16, regime="local", device=dev, seed=5 + i) for i in range(3)]
refs = [syn.encoder_reference_points(syn.SWIN_T_800x1333, torch.ones(KN, 4, 2, device=dev), dev).contiguous() for _ in sets]
cargs = [(i_["value"], i_["shapes"], i_["level_start"], i_["loc"], i_["aw"]) for i_ in sets]
z = list(zip(cargs, sets, refs))
keep = _lib.get_tuning("bwd_passes")
out = []
try:
    for passes in (1, 2, 3, 4, 8, 16):
        _lib.set_tuning(bwd_passes=passes)
        us32 = kernel_us([(lambda c=c, i_=i_, r=r: fused.backward_fusedq16(*c, i_["grad_out"], r, 2)) for c, i_, r in z])
        us16 = kernel_us([(lambda c=c, i_=i_, r=r: fused.backward_fusedq_h16(*c, i_["grad_out"], r, 2)) for c, i_, r in z])
        rec = {"kind": "bwd_passes_sweep", "passes": passes, "us_fp32_acc": us32, "us_f16_acc": us16}
        out.append(rec)
        print(json.dumps(rec), flush=True)
finally:
    _lib.set_tuning(bwd_passes=keep)
with open("gpurun_out/r2ag_passes.jsonl", "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
