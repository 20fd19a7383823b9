"""Diagnostics for the tensor-memory scatter (msda_scatter_mma*.cu): per-level relative error of grad_value against the
all-reductions kernel and the fp64 oracle on a small problem, for the first- and second-generation paths, plus timings of
the scatter at config 2's launch for every setting of bwd_mma / bwd_mma_levels / tap_share.  Run on the GPU box."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ziragroundingdino_b200 as zb  # noqa: E402
from oracle import msda_oracle as O  # noqa: E402
from ziragroundingdino_b200 import _lib, fused, synthetic as syn  # noqa: E402

dev = "cuda:0"


def small_case(shapes, N, M, Lq, seed=1):
    g = torch.Generator().manual_seed(seed)
    L, S = len(shapes), sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, 32, generator=g).bfloat16()
    loc = torch.rand(N, Lq, M, L, 4, 2, generator=g) * 1.1 - 0.05
    aw = torch.softmax(torch.randn(N, Lq, M, L * 4, generator=g), -1).view(N, Lq, M, L, 4)
    gout = torch.randn(N, Lq, M * 32, generator=g).bfloat16()
    sh = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    return value, sh, lsi, loc, aw, gout


def per_level(gv, ref, shapes):
    out, s = [], 0
    for h, w in shapes:
        a, b = gv[:, s:s + h * w].double(), ref[:, s:s + h * w]
        out.append(float((a - torch.as_tensor(b)).abs().max() / max(float(np.abs(b).max()), 1e-30)))
        s += h * w
    return out


def check():
    shapes = [(40, 60), (20, 30), (10, 15), (5, 8)]
    value, sh, lsi, loc, aw, gout = small_case(shapes, 2, 8, 300)
    o_gv, _, _ = O.c_backward(value.double().numpy(), sh.numpy(), loc.double().numpy(), aw.double().numpy(), gout.double().numpy())
    args = [t.to(dev) for t in (value, sh, lsi, loc, aw, gout)]
    for name, tun in (("reductions", dict(bwd_mma=0, bwd_mma_levels=0)), ("mma v1", dict(bwd_mma=1, bwd_mma_levels=0)),
                      ("mma v2 lv2", dict(bwd_mma=1, bwd_mma_levels=2)), ("mma v2 lv4", dict(bwd_mma=1, bwd_mma_levels=4))):
        _lib.set_tuning(bwd_mma_min_units=0, **tun)
        gv, gl, ga = zb._C.ms_deform_attn_backward(*args, 64)
        torch.cuda.synchronize()
        print("%-12s per-level rel err vs fp64: %s  finite=%s" % (name, ["%.1e" % e for e in per_level(gv.cpu(), o_gv, shapes)],
                                                                  bool(torch.isfinite(gv).all())))
    _lib.set_tuning(bwd_mma=0, bwd_mma_levels=0, bwd_mma_min_units=131072)


def timings():
    out = []
    for regime in ("local", "uniform"):
        sets = [syn.core_inputs(syn.SWIN_T_800x1333, 4, dtype=torch.bfloat16, regime=regime, device=dev, seed=5 + i) for i in range(3)]
        refs = [syn.encoder_reference_points(syn.SWIN_T_800x1333, torch.ones(4, 4, 2, device=dev), dev).contiguous() for _ in sets]
        for tun in (dict(bwd_mma=0), dict(bwd_mma=0, bwd_dots=1), dict(bwd_mma=1, bwd_mma_levels=0), dict(bwd_mma=1, bwd_mma_levels=0, bwd_dots=1),
                    dict(bwd_mma=1, bwd_mma_levels=2), dict(bwd_mma=1, bwd_mma_levels=2, bwd_dots=1), dict(bwd_mma=1, bwd_mma_levels=3, bwd_dots=1),
                    dict(bwd_mma=1, bwd_mma_levels=4, bwd_dots=1), dict(bwd_mma=0, bwd_dots=1, bwd_narrow=0),
                    dict(bwd_mma=1, bwd_mma_levels=4, bwd_dots=1, bwd_narrow=0)):
            _lib.set_tuning(**dict(dict(bwd_mma=1, bwd_mma_levels=0, tap_share=0, bwd_dots=0, bwd_narrow=1), **tun))
            fns = [(lambda i_=i_, r=r: fused.backward_fusedq16(i_["value"], i_["shapes"], i_["level_start"], i_["loc"], i_["aw"], i_["grad_out"], r, 2))
                   for i_, r in zip(sets, refs)]
            ffn = [(lambda i_=i_: zb._C.ms_deform_attn_forward(i_["value"], i_["shapes"], i_["level_start"], i_["loc"], i_["aw"], 64)) for i_ in sets]
            res = {}
            for nm, ff in (("bwd_fusedq_us", fns), ("fwd_us", ffn)):
                for f in ff:
                    f()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for i in range(12):
                    ff[i % 3]()
                b.record(); torch.cuda.synchronize()
                res[nm] = a.elapsed_time(b) * 1e3 / 12
            rec = dict(regime=regime, tuning=tun, **res)
            print(rec); out.append(rec)
    _lib.set_tuning(bwd_mma=0, bwd_mma_levels=0, tap_share=0, bwd_dots=1, bwd_narrow=1)
    with open(os.path.join(ROOT, "gpurun_out", "r2_scatter_variants.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    check()
    if "--time" in sys.argv:
        timings()
