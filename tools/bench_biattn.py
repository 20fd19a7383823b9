"""Row N4 timing at the model's size (N=4 images, S=22223 image tokens, 256 text tokens, 4 heads of 256, bf16):
the BiAttentionBlock forward / forward+backward on the tcgen05 attention core vs the two library formulations, and each
kernel of the core alone with its achieved tensor throughput.  CUDA events, median of 10 after 3 warm-ups.
Usage: python tools/bench_biattn.py [tag]   -> gpurun_out/<tag>_biattn.jsonl"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ziragroundingdino_b200 import biattn, _lib
from ziragroundingdino_b200.fuse_modules import BiAttentionBlock, BiMultiHeadAttention
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
dev = torch.device("cuda:0")
B, S, T, C, E, H = 4, 22223, 256, 256, 1024, 4
torch.manual_seed(0)
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


gv_, gl_ = torch.randn_like(v), torch.randn_like(l)
def fb_fixed():           # upstream gradients supplied: times the block alone
    ov, ol = blk(v, l, mv, ml)
    torch.autograd.backward([ov, ol], [gv_, gl_])


out = open(os.path.join(ROOT, "gpurun_out", tag + "_biattn.jsonl"), "w")
def emit(rec):
    print(json.dumps(rec)); out.write(json.dumps(rec) + "\n"); out.flush()

for frozen in (False, True):
    for prm in blk.parameters():
        prm.requires_grad_(not frozen); prm.grad = None
    for core in ("kernel", "matmul") + (() if frozen else ("sdpa",)):
        BiMultiHeadAttention.use_kernel = core == "kernel"
        BiMultiHeadAttention.use_sdpa = core == "sdpa"
        n0 = _lib.launch_count()
        fb()
        own = _lib.launch_count() - n0
        torch.cuda.reset_peak_memory_stats()
        fb()
        emit(dict(what="BiAttentionBlock N=4 S=22223 n_text=256 bf16 (v/l 256, embed 1024, 4 heads)", core=core,
                  weights="frozen (the ZiRa configuration)" if frozen else "trainable", fwd_us=t(fwd),
                  fwd_bwd_us=t(fb), fwd_bwd_given_grads_us=t(fb_fixed), own_launches_fwd_bwd=own,
                  peak_mem_mb=torch.cuda.max_memory_allocated() / 1e6))

# ---- the core's kernels alone ----
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda L: torch.randn(B, L, E, device=dev, generator=g).to(torch.bfloat16)
q, k, vv, vl, gv, gl = mk(S), mk(T), mk(S), mk(T), mk(S), mk(T)
scale = 1.0 / 16
mvp, mlp = biattn._pad_mask(mv, B, S, dev), biattn._pad_mask(ml, B, T, dev)
ns = biattn.default_splits(B, H, T, S, dev)
ns64 = biattn.default_splits(B, H, T, S, dev, tile=64)
ov, sv = biattn.pv(q, k, vl, H, scale, mlp)
ol, sl = biattn.pv(k, q, vv, H, scale, mvp, nsplit=ns)
dv, dl = biattn.rowdot(gv, ov, H), biattn.rowdot(gl, ol, H)
unit = 2.0 * B * H * S * T * 256          # one logits-sized product, FLOP
_, terms_ = biattn.ds(q, gv, vv, k, vl, gl, H, scale, mvp, mlp, sv, dv, sl, dl, want_terms=True)
legs = [
    ("pv rows online (out_v)", lambda: biattn.pv(q, k, vl, H, scale, mlp), 2),
    ("pv tokens online (out_l) + combine, nsplit=%d" % ns, lambda: biattn.pv(k, q, vv, H, scale, mvp, nsplit=ns), 2),
    ("rowdot image side", lambda: biattn.rowdot(gv, ov, H), 0),
    ("pv tokens given (d_val_l) + combine", lambda: biattn.pv(k, q, gv, H, scale, mlp, col_stat=sv, nsplit=ns), 2),
    ("pv rows given (d_val_v)", lambda: biattn.pv(q, k, gl, H, scale, mvp, col_stat=sl), 2),
    ("ds rows (d_q)", lambda: biattn.ds(q, gv, vv, k, vl, gl, H, scale, mvp, mlp, sv, dv, sl, dl), 6),
    ("ds rows (d_q) storing dS terms", lambda: biattn.ds(q, gv, vv, k, vl, gl, H, scale, mvp, mlp, sv, dv, sl, dl, want_terms=True), 6),
    ("tn (d_k from stored dS) + combine", lambda: biattn.tn(terms_, q, H, scale, T), 1),
    ("ds tokens (d_k, recompute) + combine, nsplit=%d" % ns64, lambda: biattn.ds(k, gl, vl, q, vv, gv, H, scale, mlp, mvp, sl, dl, sv, dv, nsplit=ns64), 6),
]
tot_f = tot_b = 0.0
for i, (name, fn, units) in enumerate(legs):
    us = t(fn)
    emit(dict(kernel=name, us=us, tflops=units * unit / us / 1e6 if units else None))
    if i < 2: tot_f += us
    elif "recompute" not in name and name != "ds rows (d_q)": tot_b += us
emit(dict(core_fwd_us=tot_f, core_bwd_us=tot_b, logits_product_gflop=unit / 1e9))
