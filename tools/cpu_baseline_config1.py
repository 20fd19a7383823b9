"""the reference's CPU path at config 1 (single MSDeformAttn layer, N=1, Swin-T 800x1333, fp32)
timed on this host's cores: the core function (oracle.grid_sample_core = ms_deform_attn.py:90-130) and the full module
(oracle.module_forward), forward and forward+backward, both location regimes; 2 warm-up + 5 timed, median.
TEST INFRASTRUCTURE (imports oracle/); writes gpurun_out/cpu_baseline_config1.json."""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu_encoder, msda_oracle as O  # noqa: E402
from ziragroundingdino_b200 import synthetic as syn  # noqa: E402

threads = os.cpu_count() or 1
torch.set_num_threads(threads)
model = "?"
try:
    for line in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines():
        if line.startswith("Model name"):
            model = line.split(":", 1)[1].strip()
except OSError:
    pass


def med(fn, warm=2, n=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[n // 2] * 1e3


out = {"cpu": model, "threads": threads, "torch": torch.__version__, "config": "config 1: N=1, S=Lq=22223, M=8, D=32, L=4, P=4, fp32"}
shapes = syn.SWIN_T_800x1333
sh = torch.tensor(shapes)
for regime in ("local", "uniform"):
    inp = syn.core_inputs(shapes, 1, regime=regime, device="cpu", seed=1)
    v, loc, aw, go = inp["value"], inp["loc"], inp["aw"], inp["grad_out"]

    def fwd():
        with torch.no_grad():
            O.grid_sample_core(v, sh, loc, aw)

    def fwd_bwd():
        a, b, c = v.clone().requires_grad_(True), loc.clone().requires_grad_(True), aw.clone().requires_grad_(True)
        O.grid_sample_core(a, sh, b, c).backward(go)

    out["core_%s_fwd_ms" % regime] = med(fwd)
    out["core_%s_fwd_bwd_ms" % regime] = med(fwd_bwd)
S = sum(h * w for h, w in shapes)
p = cpu_encoder.make_layer_params()
src, pos = torch.randn(1, S, 256), torch.randn(1, S, 256)
refp = syn.encoder_reference_points(shapes, torch.ones(1, 4, 2), "cpu")


def mod_fwd():
    with torch.no_grad():
        O.module_forward(p, src + pos, src, None, refp, sh, 8, 4, 4)


def mod_fwd_bwd():
    x = src.clone().requires_grad_(True)
    O.module_forward(p, x + pos, x, None, refp, sh, 8, 4, 4).square().mean().backward()


out["module_fwd_ms"] = med(mod_fwd)
out["module_fwd_bwd_ms"] = med(mod_fwd_bwd)
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_baseline_config1.json"), "w"))
