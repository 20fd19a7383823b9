"""L2 scatter probe: achievable rate of 128-byte fp32 row reductions / stores into an L2-sized buffer (the shape of the
backward kernel's grad_value traffic).  JSON lines to gpurun_out/probe_scatter.jsonl."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ziragroundingdino_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
out = open(os.path.join(ROOT, "gpurun_out", "probe_scatter.jsonl"), "w")
names = {0: "red.v4.f32 x8 lanes (128 B rows)", 1: "st.v4.f32 x8 lanes", 2: "red.f32 x32 lanes",
         3: "red.v4.f32 x4 lanes (64 B fp32 rows)", 4: "red.v4.bf16x2 x4 lanes (64 B bf16 rows)",
         5: "red.v4.f32 x16 lanes (256 B contiguous)", 6: "cp.reduce.async.bulk 128 B rows from smem"}
LANES = {0: 8, 1: 8, 2: 32, 3: 4, 4: 4, 5: 16, 6: 8}       # lanes per row (mode 6: 4 rows per warp)
BYTES = {0: 128, 1: 128, 2: 128, 3: 64, 4: 64, 5: 256, 6: 128}
for mb in (1, 23, 91, 400):
    buf = torch.zeros(mb << 18, dtype=torch.float32, device=dev)
    for mode in (0, 1, 2, 3, 4, 5, 6):
        for blocks in (148 * 4, 148 * 8, 148 * 16):
            iters = 256
            fn = lambda: L.msda_b200_probe_scatter(buf.data_ptr(), buf.numel() * 4, mode, iters, blocks, torch.cuda.current_stream().cuda_stream)
            for _ in range(2):
                assert fn() == 0
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            us = sorted(ts)[2]
            rows = blocks * 256 * iters // LANES[mode]
            rec = dict(kind="scatter_probe", buf_mb=mb, mode=names[mode], blocks=blocks, us=us, rows_per_us=rows / us,
                       gbps=rows * BYTES[mode] / us / 1e3, rows_per_clk_per_sm=rows / us / 1e3 / 1.92 / 148)
            print(rec); out.write(json.dumps(rec) + "\n")
