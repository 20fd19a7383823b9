"""GEMM timing / profiling driver: prof_gemm.py [time|run] -- value-proj, query-proj, dgrad-K384, FFN shapes at R = 4*22223."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ziragroundingdino_b200 import _lib, fused, layer_ops  # noqa: E402

dev = "cuda:0"
R = 4 * 22223
torch.manual_seed(0)
x = torch.randn(R, 256, device=dev).bfloat16()
x384 = torch.randn(R, 384, device=dev).bfloat16()
w = (torch.randn(256, 256, device=dev) * 0.05).bfloat16()
w384 = (torch.randn(256, 384, device=dev) * 0.05).bfloat16()
wq = (torch.randn(384, 256, device=dev) * 0.05).bfloat16()
w1 = (torch.randn(2048, 256, device=dev) * 0.05).bfloat16()
b = torch.randn(256, device=dev)
bq = torch.randn(384, device=dev)
b1 = torch.randn(2048, device=dev)
ref = torch.rand(R, 4, 2, device=dev)
shapes = torch.tensor([(100, 167), (50, 84), (25, 42), (13, 21)], device=dev)
h = torch.randn(R, 2048, device=dev).bfloat16()
bits = torch.empty(2048 // 32, R, dtype=torch.int32, device=dev)
acc = torch.randn(R, 256, device=dev).bfloat16()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()

cases = {
    "value_proj 256->256": lambda: fused.linear16(x, w, b),
    "query_proj 256->384 (loc+softmax)": lambda: fused.query_proj16(x, wq, bq, ref, 2, shapes, 8, 4, 4),
    "dgrad K=384 ->256": lambda: fused.linear16(x384, w384, None),
    "ffn1 256->2048 relu": lambda: layer_ops._linear_act16(x, w1, b1, relu=True),
    "ffn2 dgrad 256->2048 gated": lambda: layer_ops._linear_act16(x, w1, None, gate=h),
    "ffn1 256->2048 relu + bits": lambda: layer_ops._linear_act_bits16(x, w1, b1, relu_bits=bits),
    "ffn2 dgrad 256->2048 bit-gated": lambda: layer_ops._linear_act_bits16(x, w1, None, gate_bits=bits),
    "dgrad 256->256 accumulate": lambda: __import__("ziragroundingdino_b200.blocks", fromlist=["x"]).linear_accum16(x, w, acc),
    "cublas 256->2048 + relu": lambda: torch.relu(torch.nn.functional.linear(x, w1, b1.bfloat16())),
    "cublas 256->256": lambda: torch.nn.functional.linear(x, w, b.bfloat16()),
}


def timeit(fn, flush_l2):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        if flush_l2:
            flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


if sys.argv[1:] and sys.argv[1] == "run":
    which = sys.argv[2]
    L.msda_b200_gemm_set_staged(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    for _ in range(3):
        cases[which]()
    torch.cuda.synchronize()
else:
    for bufs, pres in ((1, 1),):
        L.msda_b200_gemm_set_store_bufs(bufs, pres)
        for name, fn in cases.items():
            print("store_bufs=%d prefer_resident=%d  %-36s cold %7.1f us   warm %7.1f us" % (bufs, pres, name, timeit(fn, True), timeit(fn, False)))
