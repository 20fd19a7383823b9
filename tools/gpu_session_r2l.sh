#!/bin/bash
# Session l: source-level stall samples of the bidirectional-attention kernels (one launch each).
mkdir -p gpurun_out
cat > /tmp/one_pv.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ziragroundingdino_b200 import biattn
dev = torch.device("cuda:0"); B, S, T, E, H = 4, 22223, 256, 1024, 4
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda L: torch.randn(B, L, E, device=dev, generator=g).to(torch.bfloat16)
q, k, vv, vl, gv, gl = mk(S), mk(T), mk(S), mk(T), mk(S), mk(T)
mvp, mlp = biattn._pad_mask(None, B, S, dev), biattn._pad_mask(None, B, T, dev)
ov, sv = biattn.pv(q, k, vl, H, 1 / 16, mlp)
ns = biattn.default_splits(B, H, T, S, dev)
ol, sl = biattn.pv(k, q, vv, H, 1 / 16, mvp, nsplit=ns)
dv, dl = biattn.rowdot(gv, ov, H), biattn.rowdot(gl, ol, H)
biattn.ds(q, gv, vv, k, vl, gl, H, 1 / 16, mvp, mlp, sv, dv, sl, dl)
torch.cuda.synchronize()
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:biattn_pv_kernel -c 1 -o gpurun_out/r2l_pv -f python /tmp/one_pv.py > gpurun_out/r2l_ncu.log 2>&1
python tools/ncu_source_lines.py gpurun_out/r2l_pv.ncu-rep > gpurun_out/r2l_pv_rows_source.txt 2>&1
rm -f gpurun_out/r2l_pv.ncu-rep
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:biattn_ds_kernel -c 1 -o gpurun_out/r2l_ds -f python /tmp/one_pv.py >> gpurun_out/r2l_ncu.log 2>&1
python tools/ncu_source_lines.py gpurun_out/r2l_ds.ncu-rep > gpurun_out/r2l_ds_rows_source.txt 2>&1
rm -f gpurun_out/r2l_ds.ncu-rep
head -30 gpurun_out/r2l_pv_rows_source.txt
