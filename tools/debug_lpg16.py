"""Where does the 16-bit D=64 / D=128 backward (16 lanes per unit) differ from the oracle?  Prints error location stats."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ziragroundingdino_b200 as zb
from oracle import msda_oracle as O
from ziragroundingdino_b200 import _lib
dev = "cuda:0"
def mk(shapes, N, M, D, Lq, seed, dtype, lo, hi):
    g = torch.Generator().manual_seed(seed)
    L, S = len(shapes), sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g).to(dtype)
    loc = torch.rand(N, Lq, M, L, 4, 2, generator=g) * (hi - lo) + lo
    aw = torch.softmax(torch.randn(N, Lq, M, L * 4, generator=g), -1).view(N, Lq, M, L, 4)
    gout = torch.randn(N, Lq, M * D, generator=g).to(dtype)
    sh = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    return value, sh, lsi, loc, aw, gout
shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
for (D, M, Lq, lo, hi) in ((64, 4, 333, -0.2, 1.2), (64, 4, 333, -0.05, 1.05), (64, 8, 200, -0.2, 1.2), (64, 4, 336, -0.2, 1.2), (64, 4, 64, 0.1, 0.9), (128, 2, 333, -0.2, 1.2)):
    for narrow in (1, 0):
        _lib.set_tuning(bwd_dots=0, bwd_mma=0, bwd_narrow=narrow)
        value, sh, lsi, loc, aw, gout = mk(shapes, 2, M, D, Lq, 80 + D, torch.bfloat16, lo, hi)
        gv, gl, ga = zb._C.ms_deform_attn_backward(*[t.to(dev) for t in (value, sh, lsi, loc, aw, gout)], 64)
        ogv, ogl, oga = O.c_backward(value.double().numpy(), sh.numpy(), loc.double().numpy(), aw.double().numpy(), gout.double().numpy())
        e = np.abs(gl.cpu().numpy() - ogl); ea = np.abs(ga.cpu().numpy() - oga)
        bad = np.argwhere(e > 1e-3 * np.abs(ogl).max())
        print("D=%d M=%d Lq=%d loc[%.2f,%.2f] narrow=%d: gl rel %.2e ga rel %.2e gv rel %.2e  #bad gl=%d" % (
            D, M, Lq, lo, hi, narrow, e.max() / np.abs(ogl).max(), ea.max() / np.abs(oga).max(),
            np.abs(gv.float().cpu().numpy() - ogv).max() / np.abs(ogv).max(), len(bad)))
        if len(bad):
            print("   first bad (n,q,m,l,p,xy):", bad[:6].tolist(), " q values:", sorted(set(bad[:, 1].tolist()))[:12], " m:", sorted(set(bad[:, 2].tolist())),
                  " l:", sorted(set(bad[:, 3].tolist())))
            n, q, m, l, p, xy = bad[0]
            print("   got %.5f want %.5f loc=(%.4f, %.4f)" % (gl[n, q, m, l, p, xy].item(), ogl[n, q, m, l, p, xy], loc[n, q, m, l, p, 0].item(), loc[n, q, m, l, p, 1].item()))
_lib.set_tuning(bwd_narrow=1)
