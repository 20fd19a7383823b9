"""HBM write-only / read-only / copy bandwidth with library kernels, for the GEMM rooflines (the FFN GEMMs write 364 MB)."""
import json, os, sys
import torch
dev = torch.device("cuda:0")
def t(fn, iters=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters
out = {}
for mb in (364, 1024, 4096):
    n = mb * (1 << 20) // 2
    x = torch.empty(n, dtype=torch.bfloat16, device=dev); y = torch.empty_like(x)
    us = t(lambda: x.fill_(1.0)); out["fill_%dMB" % mb] = dict(us=us, gbps=n * 2 / us / 1e3)
    us = t(lambda: x.zero_()); out["memset_%dMB" % mb] = dict(us=us, gbps=n * 2 / us / 1e3)
    us = t(lambda: y.copy_(x)); out["copy_%dMB" % mb] = dict(us=us, gbps=n * 4 / us / 1e3)
    us = t(lambda: torch.linalg.vector_norm(x.view(torch.int16)[: n // 2 * 2].view(torch.bfloat16), 2, dtype=torch.float32)); out["read_%dMB" % mb] = dict(us=us, gbps=n * 2 / us / 1e3)
    us = t(lambda: torch.relu_(x)); out["relu_inplace_%dMB" % mb] = dict(us=us, gbps=n * 4 / us / 1e3)
    del x, y
for k, v in out.items(): print(k, "%.1f us  %.0f GB/s" % (v["us"], v["gbps"]))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "probe_hbm.json"), "w"))
