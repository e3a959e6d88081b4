"""One (C, k, d) fused ResBlock pair at the vocoder stage size, a few launches -- the ncu target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import _native as N
dev = torch.device("cuda", 0)
C, k, d = int(os.environ.get("C", 16)), int(os.environ.get("K", 3)), int(os.environ.get("D", 1))
B = int(os.environ.get("B", 32)); T = int(os.environ.get("T", 256000 * 128 // C))
pc = N.PREC_F16
x = torch.randn(B, T, C, device=dev)
pa = N.split_nwc(torch.where(x >= 0, x, x * 0.1), pc); del x
out = torch.empty_like(pa)
w = torch.randn(C, k * C, device=dev) / (k * C) ** 0.5
s = N.pow2_scale(w); wp = N.pack_weight(w, pc, s); b = torch.zeros(C, device=dev)
for _ in range(int(os.environ.get("N", 3))):
    N.respair(pa, wp, wp, b, b, B, T, C, k, d, k, out_planes=out, w1_inv_scale=1 / s, w2_inv_scale=1 / s)
torch.cuda.synchronize()
print("done")
