"""MMA-issuer wait breakdown of the CTA-pair chained FFN (build with MSDA_NVCC_EXTRA=-DFFN2_TRACE)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MSDA_B200_FFN_PAIR"] = "1"
from ziragroundingdino_b200 import blocks, _lib
dev = "cuda:0"
R, C, F = 88892, 256, 2048
g = torch.Generator().manual_seed(1)
x = torch.randn(R, C, generator=g).bfloat16().to(dev)
w1 = (torch.randn(F, C, generator=g) * 0.06).bfloat16().to(dev); w2 = (torch.randn(C, F, generator=g) * 0.02).bfloat16().to(dev)
b1 = torch.randn(F, generator=g).to(dev) * 0.1; b2 = torch.randn(C, generator=g).to(dev) * 0.1
bits = torch.zeros((F // 32, R), dtype=torch.int32, device=dev)
L = _lib.lib(); L.msda_ffn_chain2_set_trace.argtypes = [ctypes.c_void_p]
names = ["x_full", "a1_free", "full(G1)", "a2_free", "h_full", "full(G2)", "TOTAL", "-"]
for pair in (True,):
    blocks.FFN_PAIR = pair
    blocks.ffn_chain_fwd16(x, w1, b1, w2, b2, bits); torch.cuda.synchronize()
    tr = torch.zeros(2 * 74 * 8, dtype=torch.int64, device=dev)
    L.msda_ffn_chain2_set_trace(tr.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); blocks.ffn_chain_fwd16(x, w1, b1, w2, b2, bits); b.record(); torch.cuda.synchronize()
    L.msda_ffn_chain2_set_trace(0)
    t = tr[:592].view(74, 8).double().mean(0).cpu().tolist()
    e = tr[592:].view(74, 8).double().mean(0).cpu().tolist()
    print("pair fwd %.1f us :" % (a.elapsed_time(b) * 1e3), "  ".join("%s=%.0f" % (n, v) for n, v in zip(names, t) if n != "-"))
    print("  mid warp 2 (leader CTA):", "  ".join("%s=%.0f" % (n, v) for n, v in zip(["wait a1_full", "wait h_free", "tmem_ld", "math+sts", "fence+arrive", "final", "TOTAL"], e)))
