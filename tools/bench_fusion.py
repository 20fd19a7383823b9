"""Row N4 timing: BiAttentionBlock at config-2 size (N=4 images, S=22223 image tokens, 256 text tokens, bf16), fwd and
fwd+bwd, CUDA events, median of 10.  Prints the bytes the reference's formulation would materialise for comparison."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ziragroundingdino_b200.fuse_modules import BiAttentionBlock
dev = torch.device("cuda:0")
B, S, T, C, E, H = 4, 22223, 256, 256, 1024, 4
blk = BiAttentionBlock(C, C, E, H, dropout=0.0, drop_path=0.0).to(dev).to(torch.bfloat16)
v = torch.randn(B, S, C, device=dev).to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B, T, C, device=dev).to(torch.bfloat16).requires_grad_(True)
mv = torch.zeros(B, S, dtype=torch.bool, device=dev); mv[1, -3000:] = True
ml = torch.zeros(B, T, dtype=torch.bool, device=dev); ml[:, -56:] = True
def t(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[n // 2]
def fwd():
    with torch.no_grad():
        blk(v, l, mv, ml)
def fb():
    ov, ol = blk(v, l, mv, ml)
    (ov.float().square().mean() + ol.float().square().mean()).backward()
from ziragroundingdino_b200.fuse_modules import BiMultiHeadAttention
recs = []
for sdpa in (False, True):
    BiMultiHeadAttention.use_sdpa = sdpa
    recs.append(dict(what="BiAttentionBlock N=4 S=22223 n_text=256 bf16 (v/l 256, embed 1024, 4 heads)", use_sdpa=sdpa,
                     fwd_us=t(fwd), fwd_bwd_us=t(fb), reference_attention_matrix_mb=B * H * S * T * 4 / 1e6 * 2))
    print(json.dumps(recs[-1]))
json.dump(recs, open(os.path.join(ROOT, "gpurun_out", "bench_fusion.json"), "w"))
