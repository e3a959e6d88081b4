"""Device time of the fused ResBlock pair kernel per (C, k, dilation) at the NSF-HiFiGAN stage sizes (B=32, T=4000
frames, hop 512): rows per item = 256000 * 128 / C."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import _native as N
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 32))
only = os.environ.get("C")
pc = N.PREC_F16
for C in (128, 64, 32):
    if only and int(only) != C:
        continue
    if os.environ.get("C32_64") and C == 128:
        continue
    T = 256000 * 128 // C
    x = torch.randn(B, T, C, device=dev)
    pa = N.split_nwc(torch.where(x >= 0, x, x * 0.1), pc)
    del x
    out = torch.empty_like(pa)
    xs = torch.empty((B, T, C), dtype=torch.float32, device=dev)
    for k in (3, 7, 11):
        w = torch.randn(C, k * C, device=dev) / (k * C) ** 0.5
        s = N.pow2_scale(w)
        wp = N.pack_weight(w, pc, s)
        b = torch.zeros(C, device=dev)
        for d in (1, 3, 5):
            for mode in ("planes",):
                kw = dict(out_planes=out)
                N.respair(pa, wp, wp, b, b, B, T, C, k, d, k, w1_inv_scale=1 / s, w2_inv_scale=1 / s, **kw)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    N.respair(pa, wp, wp, b, b, B, T, C, k, d, k, w1_inv_scale=1 / s, w2_inv_scale=1 / s, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                fl = 2.0 * B * T * C * C * 2 * k
                by = B * T * C * (8 if mode == "planes" else 12)
                print(f"C={C:4d} k={k:2d} d={d} {mode:6s} {ms:8.3f} ms  {fl/ms/1e9:7.1f} TFLOP/s alg  {by/ms/1e6:7.0f} GB/s alg", flush=True)
