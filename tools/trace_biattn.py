"""Pipeline-wait breakdown of biattn_pv_kernel (debug counters, cycles summed per CTA): where the MMA issuer and one
mid-stage warp spend their time, rows and tokens orientation at the model's size.
Needs a traced build: MSDA_NVCC_EXTRA=-DBIA_TRACE bash ziragroundingdino_b200/csrc/build.sh (the default build has no counters)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ziragroundingdino_b200 import biattn, _lib
dev = torch.device("cuda:0")
B, S, T, E, H = 4, 22223, 256, 1024, 4
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda L: torch.randn(B, L, E, device=dev, generator=g).to(torch.bfloat16)
q, k, vv, vl = mk(S), mk(T), mk(S), mk(T)
mvp, mlp = biattn._pad_mask(None, B, S, dev), biattn._pad_mask(None, B, T, dev)
L = _lib.lib()
L.msda_biattn_set_trace.argtypes = [ctypes.c_void_p]
ns = biattn.default_splits(B, H, T, S, dev)
names_m = ["x_full", "a1_free", "full1(B)", "a2_free", "h_full(P)", "full2(X)", "TOTAL", "-"]
names_e = ["a1_full(S)", "pair_bar", "h_free", "a2_full", "final", "tmem_ld", "TOTAL", "exp_pack"]
for what, fn in (("rows online", lambda: biattn.pv(q, k, vl, H, 1 / 16, mlp)),
                 ("tokens online nsplit=%d" % ns, lambda: biattn.pv(k, q, vv, H, 1 / 16, mvp, nsplit=ns))):
    fn(); torch.cuda.synchronize()
    tr = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    L.msda_biattn_set_trace(tr.data_ptr())
    fn(); torch.cuda.synchronize()
    L.msda_biattn_set_trace(0)
    t = tr.view(148, 16).double().mean(0).cpu().tolist()
    print("==", what)
    print("  MMA issuer :", "  ".join("%s=%.0f" % (n, v) for n, v in zip(names_m, t[:8]) if n != "-"))
    print("  mid warp 2 :", "  ".join("%s=%.0f" % (n, v) for n, v in zip(names_e, t[8:]) if n != "-"))
