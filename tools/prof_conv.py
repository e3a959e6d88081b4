"""One vocoder-style conv through fd_conv_cl_fwd (for ncu captures / quick timing of a single kernel shape).

    python tools/prof_conv.py --cin 16 --n 16 --taps 11 --rows 2048000 --kind c2 [--precision f16]
kind: c1 = bias + LeakyReLU -> planes;  c2 = + fp32 residual, fp32 master out and LeakyReLU planes out.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cin", type=int, default=16)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--taps", type=int, default=11)
    ap.add_argument("--dil", type=int, default=1)
    ap.add_argument("--rows", type=int, default=2048000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--kind", default="c2")
    ap.add_argument("--precision", default="f16")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from fish_diffusion_b200 import _native as N
    dev = torch.device("cuda:0")
    B, T, Ci, Nn = args.batch, args.rows, args.cin, args.n
    pc, mma = N.prec_code(args.precision), N.mma_code(args.precision)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((B, T, Ci), device=dev, generator=g)
    xp = N.split_nwc(x, pc)
    del x
    shifts = [(j - args.taps // 2) * args.dil for j in range(args.taps)]
    w = torch.randn((Nn, args.taps * Ci), device=dev, generator=g) / (args.taps * Ci) ** 0.5
    s = N.pow2_scale(w)
    wp = N.pack_weight(w, pc, s)
    bias = torch.randn(Nn, device=dev, generator=g)
    outp = torch.empty((2, B, T, Nn), dtype=torch.int16, device=dev)
    kw = dict(bias=bias, out_planes=outp, w_inv_scale=1.0 / s, act=N.ACT_LRELU, act_slope=0.1, prec=mma,
              backend=N.BACKEND_TC)
    if args.kind == "c2":
        kw.update(res_f32=torch.randn((B, T, Nn), device=dev, generator=g),
                  out_f32=torch.empty((B, T, Nn), device=dev))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
    N.conv_cl(xp, wp, B, T, Ci, Nn, shifts, **kw)
    ev[0].record()
    for i in range(args.reps):
        N.conv_cl(xp, wp, B, T, Ci, Nn, shifts, **kw)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps)]
    byts = B * T * (Ci * 4 + Nn * 4 + (Nn * 8 if args.kind == "c2" else 0))
    print(f"cin={Ci} n={Nn} taps={args.taps} kind={args.kind} {args.precision}: {min(ms):.3f} ms  "
          f"{byts / min(ms) / 1e6:.0f} GB/s algorithmic  {2 * B * T * Ci * Nn * args.taps / min(ms) / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
