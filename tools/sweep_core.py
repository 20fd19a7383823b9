"""GPU sweep of the core gather/scatter kernels: tuning variants, the L2 gather probe and the
reference's own CUDA op (oracle/_ref) on the same inputs.  Writes JSON lines to gpurun_out/."""
import ctypes
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ziragroundingdino_b200 as zb  # noqa: E402
from ziragroundingdino_b200 import _lib, synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, flush=True):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(f, **kw):
    f.write(json.dumps(kw) + "\n"); f.flush()
    print(kw)


def probe(f):
    L = _lib.lib()
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    for mb in (11, 22, 45, 91, 400):
        buf = torch.randint(0, 255, (mb << 20,), dtype=torch.uint8, device=dev)
        for seg in (64, 128, 512):
            for blocks in (148 * 4, 148 * 8, 148 * 16):
                iters = 512
                fn = lambda: L.msda_b200_probe_gather(buf.data_ptr(), buf.numel(), seg, iters, blocks, sink.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream)
                med, best = timeit(fn, iters=5, flush=False)
                bytes_ = blocks * 256 * 16 * iters
                emit(f, kind="probe", buf_mb=mb, seg=seg, blocks=blocks, us=med, gbps=bytes_ / med / 1e3,
                     gbps_best=bytes_ / best / 1e3)
        del buf


def core(f, name, shapes, N, dtype, regime, Lq=None, full=True):
    inp = syn.core_inputs(shapes, N, dtype=dtype, regime=regime, Lq=Lq, device=dev)
    Nn, S, M, D, L, Lq_, P = inp["dims"]
    ab = syn.algorithmic_bytes(Nn, S, M, D, L, Lq_, P, inp["value"].element_size())
    args = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
    fwd = lambda: zb._C.ms_deform_attn_forward(*args, 64)
    bwd = lambda: zb._C.ms_deform_attn_backward(*args, inp["grad_out"], 64)
    keep = {k: _lib.get_tuning(k) for k in ("fwd_sample_batch", "fwd_q_fast", "fwd_passes", "bwd_q_fast", "bwd_passes")}
    combos = list(itertools.product((1, 2, 4), (0, 1), (1, 4, 16))) if full else [(keep["fwd_sample_batch"], keep["fwd_q_fast"], keep["fwd_passes"])]
    for sb, qf, ps in combos:
        _lib.set_tuning(fwd_sample_batch=sb, fwd_q_fast=qf, fwd_passes=ps)
        for flush in (True, False):
            med, best = timeit(fwd, flush=flush)
            emit(f, kind="fwd", case=name, sb=sb, qf=qf, passes=ps, flush=flush, us=med, us_best=best,
                 l2_gbps=ab["fwd_l2"] / med / 1e3, hbm_gbps=ab["fwd_hbm"] / med / 1e3)
    _lib.set_tuning(**keep)
    combos = list(itertools.product((0, 1), (1, 4, 16))) if full else [(keep["bwd_q_fast"], keep["bwd_passes"])]
    for qf, ps in combos:
        _lib.set_tuning(bwd_q_fast=qf, bwd_passes=ps)
        for flush in (True, False):
            med, best = timeit(bwd, flush=flush)
            emit(f, kind="bwd", case=name, qf=qf, passes=ps, flush=flush, us=med, us_best=best,
                 l2_gbps=ab["bwd_l2"] / med / 1e3, hbm_gbps=ab["bwd_hbm"] / med / 1e3)
    _lib.set_tuning(**keep)
    if dtype != torch.float32:
        keep_n = _lib.get_tuning("bwd_narrow")
        for nw in (0, 1):
            _lib.set_tuning(bwd_narrow=nw)
            med, best = timeit(bwd, flush=True)
            emit(f, kind="bwd_narrow", case=name, narrow=nw, us=med, us_best=best, l2_gbps=ab["bwd_l2"] / med / 1e3)
        _lib.set_tuning(bwd_narrow=keep_n)
    if dtype == torch.float32:
        from oracle import build_ref
        ref = build_ref.load()
        if ref is not None:
            rf = lambda: ref.ms_deform_attn_forward(*args, 64)
            rb = lambda: ref.ms_deform_attn_backward(*args, inp["grad_out"], 64)
            for flush in (True, False):
                med, best = timeit(rf, flush=flush)
                emit(f, kind="ref_fwd", case=name, flush=flush, us=med, us_best=best, l2_gbps=ab["fwd_l2"] / med / 1e3)
                med, best = timeit(rb, flush=flush)
                emit(f, kind="ref_bwd", case=name, flush=flush, us=med, us_best=best, l2_gbps=ab["bwd_l2"] / med / 1e3)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    with open(os.path.join(OUT, "sweep_%s.jsonl" % tag), "w") as f:
        emit(f, kind="env", gpu=torch.cuda.get_device_name(0), sms=torch.cuda.get_device_properties(0).multi_processor_count)
        if "--no-probe" not in sys.argv:
            probe(f)
        quick = "--quick" in sys.argv
        core(f, "c1_f32_local_N1", syn.SWIN_T_800x1333, 1, torch.float32, "local", full=not quick)
        core(f, "c1_f32_uniform_N1", syn.SWIN_T_800x1333, 1, torch.float32, "uniform", full=False)
        core(f, "c2_bf16_local_N4", syn.SWIN_T_800x1333, 4, torch.bfloat16, "local", full=not quick)
        core(f, "c2_bf16_uniform_N4", syn.SWIN_T_800x1333, 4, torch.bfloat16, "uniform", full=False)
        core(f, "c2_f32_local_N4", syn.SWIN_T_800x1333, 4, torch.float32, "local", full=False)
        core(f, "dec_bf16_local_N4", syn.SWIN_T_800x1333, 4, torch.bfloat16, "local", Lq=900, full=False)
        core(f, "c5_bf16_local_N2_s8", syn.SWIN_B_1024x1800_S8, 2, torch.bfloat16, "local", full=False)
        core(f, "c5_bf16_local_N2_s4", syn.SWIN_B_1024x1800_S4, 2, torch.bfloat16, "local", full=False)


if __name__ == "__main__":
    main()
