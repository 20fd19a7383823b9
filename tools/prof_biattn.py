"""Kernel-level breakdown of one BiAttentionBlock forward+backward at the model's size (torch.profiler, CUDA time)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ziragroundingdino_b200.fuse_modules import BiAttentionBlock, BiMultiHeadAttention
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
B, S, T, C, E, H = 4, 22223, 256, 256, 1024, 4
frozen = len(sys.argv) > 1 and sys.argv[1] == "frozen"
torch.manual_seed(0)
blk = BiAttentionBlock(C, C, E, H, dropout=0.0, drop_path=0.0).to(dev).to(torch.bfloat16)
if frozen:
    for p in blk.parameters(): p.requires_grad_(False)
v = torch.randn(B, S, C, device=dev).to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B, T, C, device=dev).to(torch.bfloat16).requires_grad_(True)
mv = torch.zeros(B, S, dtype=torch.bool, device=dev); mv[1, -3000:] = True
ml = torch.zeros(B, T, dtype=torch.bool, device=dev); ml[:, -56:] = True
gv, gl = torch.randn_like(v), torch.randn_like(l)
def fb():
    ov, ol = blk(v, l, mv, ml)
    torch.autograd.backward([ov, ol], [gv, gl])
for _ in range(3): fb()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
