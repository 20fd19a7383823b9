import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ziragroundingdino_b200 as zb
from oracle import msda_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(86, 32, dtype=torch.float64); w = torch.randn(48, 32, dtype=torch.float64); b = torch.randn(48, dtype=torch.float64)
print("fp64 linear gpu-vs-cpu", (torch.nn.functional.linear(x.to(dev), w.to(dev), b.to(dev)).cpu() - torch.nn.functional.linear(x, w, b)).abs().max().item())
print("fp64 softmax gpu-vs-cpu", (x.to(dev).softmax(-1).cpu() - x.softmax(-1)).abs().max().item())
print("allow_tf32", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
g = load_golden("core_tiny_f64")
t = {k: torch.from_numpy(v) for k, v in g.items()}
sh = t["shapes"]; lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
out = zb._C.ms_deform_attn_forward(t["value"].to(dev), sh.to(dev), lsi.to(dev), t["loc"].to(dev), t["aw"].to(dev), 64)
print("core f64 vs golden", (out.cpu() - t["out_f64"]).abs().max().item())
g = load_golden("module_enc")
C, M, L, P, bf = (int(v) for v in g["cfg"])
m = zb.MultiScaleDeformableAttention(C, M, L, P, batch_first=True)
m.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
m = m.to(dev).double()
q = torch.from_numpy(g["query"]).to(dev); v = torch.from_numpy(g["value"]).to(dev)
sh = torch.from_numpy(g["shapes"]).to(dev); lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
refp = torch.from_numpy(g["reference_points"]).to(dev); mask = torch.from_numpy(g["mask"]).to(dev)
for k, p in m.state_dict().items():
    print(k, p.dtype, (p.cpu() - torch.from_numpy(g["param." + k])).abs().max().item())
out = m(query=q, value=v, key_padding_mask=mask, reference_points=refp, spatial_shapes=sh, level_start_index=lsi)
print("module f64 vs golden", (out.cpu() - torch.from_numpy(g["out"])).abs().max().item())
