"""fwd_passes sweep (consecutive unit tiles per CTA) of the forward gather at config 2's launch (4 and 8 images)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import ziragroundingdino_b200 as zb  # noqa: E402
from ziragroundingdino_b200 import _lib, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")


def kernel_us(fns, reps=24):
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


keep = _lib.get_tuning("fwd_passes")
out = []
try:
    for KN in (4, 8):
        sets = [syn.core_inputs(syn.SWIN_T_800x1333, KN, dtype=torch.bfloat16, regime="local", device=dev, seed=5 + i) for i in range(3)]
        cargs = [(i_["value"], i_["shapes"], i_["level_start"], i_["loc"], i_["aw"]) for i_ in sets]
        for passes in (1, 2, 3, 4, 8):
            _lib.set_tuning(fwd_passes=passes)
            us = kernel_us([(lambda c=c: zb._C.ms_deform_attn_forward(*c, 64)) for c in cargs])
            rec = {"kind": "fwd_passes_sweep", "images": KN, "passes": passes, "us_fwd": us}
            out.append(rec)
            print(json.dumps(rec), flush=True)
        del sets, cargs
finally:
    _lib.set_tuning(fwd_passes=keep)
with open("gpurun_out/r2ak_fwd_passes.jsonl", "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
