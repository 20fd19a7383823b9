"""Module-level timings for BASELINE.json configs 1, 3 and 5 (config 2 / 4 are bench.py).  JSON lines to
gpurun_out/configs_<tag>.jsonl.  CUDA-event timing, median of 10 after 3 warm-ups, L2 flushed between iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ziragroundingdino_b200 as zb  # noqa: E402
from ziragroundingdino_b200 import _lib, encoder, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def make_module(L, dtype, seed=0):
    torch.manual_seed(seed)
    m = zb.MultiScaleDeformableAttention(256, 8, L, 4, batch_first=True)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.02)
    return m.to(dev).to(dtype)


def config1(f):
    shapes = syn.SWIN_T_800x1333
    S = sum(h * w for h, w in shapes)
    sh, lsi = syn.level_tensors(shapes, dev)
    torch.manual_seed(1)
    for dtype in (torch.float32, torch.bfloat16):
        m = make_module(4, dtype, 1)
        src = torch.randn(1, S, 256, device=dev).to(dtype).requires_grad_(True)
        pos = torch.randn(1, S, 256, device=dev).to(dtype)
        refp = encoder.get_reference_points(shapes, torch.ones(1, 4, 2, device=dev), dev)
        mask = torch.zeros(1, S, dtype=torch.bool, device=dev)
        kw = dict(reference_points=refp, spatial_shapes=sh, level_start_index=lsi, key_padding_mask=mask)
        fwd = lambda: m(query=src + pos, value=src, **kw)
        with torch.no_grad():
            t_f = timeit(fwd)

        def fb():
            src.grad = None
            fwd().float().square().mean().backward()
        t_fb = timeit(fb)
        for p_ in m.parameters():          # the ZiRa configuration: the module's own linears are frozen (no weight-gradient GEMMs)
            p_.requires_grad_(False)
        n0 = _lib.launch_count()
        fb()
        launches = _lib.launch_count() - n0
        t_fb_frozen = timeit(fb)
        rec = dict(config=1, what="single MSDeformAttn module, N=1, Swin-T 800x1333, Lq=S=22223", dtype=str(dtype),
                   fwd_us=t_f, fwd_bwd_us=t_fb, fwd_bwd_frozen_us=t_fb_frozen, own_launches_fwd_bwd_frozen=launches,
                   fused=bool(m._use_fused(src, refp) or m._use_fused32(src, refp)))
        print(rec); f.write(json.dumps(rec) + "\n"); f.flush()


def config3(f):
    """Decoder cross-attention x6: 900 queries, 4-d reference boxes, shared memory, per-layer value_proj; ZiRa
    branches on value_proj/output_proj: train-mode un-merged vs merged (after __rep__) vs eval."""
    shapes = syn.SWIN_T_800x1333
    S = sum(h * w for h, w in shapes)
    sh, lsi = syn.level_tensors(shapes, dev)
    N, Lq = 4, 900
    for dtype in (torch.float32, torch.bfloat16):
        torch.manual_seed(3)
        layers = [make_module(4, dtype, 30 + i).add_zira_branches() for i in range(6)]
        with torch.no_grad():
            for m in layers:
                for ad in (m.value_proj_adapter, m.output_proj_adapter):
                    ad.weight.normal_(0, 1e-3); ad.freeze_linear.weight.normal_(0, 1e-3)
        memory = torch.randn(N, S, 256, device=dev).to(dtype)
        tgt = torch.randn(N, Lq, 256, device=dev).to(dtype)
        mask, valid = encoder.padded_batch_masks(shapes, N, dev, generator=torch.Generator().manual_seed(3))
        box = torch.cat([torch.rand(N, Lq, 1, 2, device=dev) * 0.8 + 0.1, torch.rand(N, Lq, 1, 2, device=dev) * 0.45 + 0.05], -1)
        refp = (box * torch.cat([valid, valid], -1)[:, None]).contiguous()

        def run():
            x = tgt
            for m in layers:
                x = x + m(query=x, value=memory, reference_points=refp, spatial_shapes=sh, level_start_index=lsi,
                          key_padding_mask=mask)
            return x
        res = {}
        for m in layers:
            m.train()
        with torch.no_grad():
            y_train = run(); res["train_unmerged_us"] = timeit(run)
        for m in layers:
            m.eval()
        with torch.no_grad():
            res["eval_unmerged_us"] = timeit(run)
        for m in layers:
            zb.merge_all(m)
        with torch.no_grad():
            y_merged = run(); res["eval_merged_us"] = timeit(run)
        for m in layers:
            m.train()
        with torch.no_grad():
            res["train_merged_us"] = timeit(run)
        diff = (y_merged.float() - y_train.float()).abs().max().item()
        rec = dict(config=3, what="decoder cross-attn x6, N=4, Lq=900, ZiRa on value/output proj", dtype=str(dtype),
                   merged_vs_unmerged_max_abs=diff, out_max_abs=y_train.float().abs().max().item(), **res)
        print(rec); f.write(json.dumps(rec) + "\n"); f.flush()


def config5(f):
    for name, shapes in (("s8", syn.SWIN_B_1024x1800_S8), ("s4", syn.SWIN_B_1024x1800_S4)):
        S = sum(h * w for h, w in shapes)
        sh, lsi = syn.level_tensors(shapes, dev)
        N, dtype = 2, torch.bfloat16
        m = make_module(5, dtype, 5)
        torch.manual_seed(5)
        src = torch.randn(N, S, 256, device=dev).to(dtype).requires_grad_(True)
        refp = encoder.get_reference_points(shapes, torch.ones(N, 5, 2, device=dev), dev)
        kw = dict(reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
        fwd = lambda: m(query=src, value=src, **kw)
        with torch.no_grad():
            t_f = timeit(fwd, 5)

        def fb():
            src.grad = None
            fwd().float().square().mean().backward()
        t_fb = timeit(fb, 5)
        inp = syn.core_inputs(shapes, N, dtype=dtype, regime="local", device=dev, seed=5)
        args = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
        c_f = timeit(lambda: zb._C.ms_deform_attn_forward(*args, 64), 5)
        c_b = timeit(lambda: zb._C.ms_deform_attn_backward(*args, inp["grad_out"], 64), 5)
        rec = dict(config=5, what="Swin-B 1024x1800, 5 levels (%s reading), N=2, bf16, encoder self-attn" % name, S=S,
                   module_fwd_us=t_f, module_fwd_bwd_us=t_fb, core_fwd_us=c_f, core_bwd_us=c_b,
                   value_map_mb=N * S * 256 * 2 / 1e6, fused=m._use_fused(src, refp))
        print(rec); f.write(json.dumps(rec) + "\n"); f.flush()


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs_%s.jsonl" % tag), "w") as f:
        config1(f); config3(f); config5(f)
