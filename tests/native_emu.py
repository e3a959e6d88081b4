"""TEST INFRASTRUCTURE ONLY -- a torch (CPU) emulation of the handful of C-ABI primitives that
fish_diffusion_b200/vocoder_train.py composes (split planes, tap-GEMM with the linear epilogue, direct weight gradient,
LeakyReLU backward, column sums).

Why it exists: there is no GPU in the build container, and the product has no CPU path (it raises on CPU tensors).  The
HOST-side logic of the training nodes -- which tap shifts, which transposed packs, the time-fold of narrow weight
gradients and its adjoint, the polyphase ConvTranspose algebra, gradient scaling -- is plain Python that can be checked
against torch autograd / the unmodified reference on the CPU if the device primitives are swapped for emulations with the
semantics documented in include/fishdiff_b200.h.  `emulated_native()` monkeypatches fish_diffusion_b200._native for the
duration of a test; nothing under fish_diffusion_b200/ imports this file and the `-m gpu` tests never use it (they run
the real kernels and compare with golden vectors of the reference).

Planes are emulated faithfully: int16 [2, ...] holding the fp16 hi / lo halves of the value (22-bit mantissa), so the
tolerances seen here are those of the real storage format; GEMMs accumulate in float64.
"""
import contextlib

import torch

from fish_diffusion_b200 import _native as N


def enc(v: torch.Tensor) -> torch.Tensor:
    v = v.to(torch.float32).clamp(-65504.0, 65504.0)
    hi = v.to(torch.float16)
    lo = (v - hi.to(torch.float32)).to(torch.float16)
    return torch.stack([hi.view(torch.int16), lo.view(torch.int16)], dim=0).contiguous()


def dec(p: torch.Tensor) -> torch.Tensor:
    return p[0].view(torch.float16).to(torch.float64) + p[1].view(torch.float16).to(torch.float64)


def _shift_rows(a, s):
    """a [B,T,C] -> rows t + s (zeros outside [0,T))"""
    B, T, C = a.shape
    out = torch.zeros_like(a)
    lo, hi = max(0, -s), min(T, T - s)
    if hi > lo:
        out[:, lo:hi] = a[:, lo + s:hi + s]
    return out


def _act(v, act, slope):
    if act == N.ACT_RELU:
        return v.clamp(min=0)
    if act == N.ACT_LRELU:
        return torch.where(v > 0, v, v * slope)
    return v


def split_nwc(x, prec, mask=None, scale=1.0, out=None):
    p = enc(x * scale)
    if mask is not None:
        p[:, mask.bool()] = 0
    if out is not None:
        out.copy_(p)
        return out
    return p


def pack_weight(w2d, prec, scale):
    return enc(w2d.detach() * scale)


def mrf_finish(ins, out, *, in_slope=0.1, scale=1.0, out_slope=0.1, prec=0):
    tot = 0
    for t in ins:
        v = dec(t)
        tot = tot + torch.where(v < 0, v / in_slope, v)
    v = tot * scale
    out.copy_(enc(torch.where(v > 0, v, v * out_slope)).view_as(out))


def conv_cl(in_planes, w_planes, B, T, Cin, Nn, shifts, *, bias=None, addend=None, res_f32=None, res_planes=None,
            row_mask=None, out_f32=None, out_planes=None, w_inv_scale=1.0, post_scale=1.0, planes_scale=1.0,
            act=0, act_slope=0.0, out_accum=False, prec=0, backend=0):
    a = dec(in_planes).reshape(B, T, Cin)
    w = dec(w_planes).reshape(Nn, len(shifts) * Cin)
    y = torch.zeros((B, T, Nn), dtype=torch.float64)
    for j, s in enumerate(shifts):
        y += _shift_rows(a, int(s)) @ w[:, j * Cin:(j + 1) * Cin].T
    y = y * w_inv_scale
    if bias is not None:
        y = y + bias.to(torch.float64)
    if addend is not None:
        y = y + addend.to(torch.float64)
    if res_f32 is not None:
        y = y + res_f32.to(torch.float64)
    if res_planes is not None:
        y = y + dec(res_planes).reshape(B, T, Nn)
    y = y * post_scale
    v = y
    if out_f32 is not None:
        if out_accum:
            v = out_f32.to(torch.float64) + y
        out_f32.copy_(v.to(torch.float32).view_as(out_f32))
    if out_planes is not None:
        out_planes.copy_(enc(_act(v * planes_scale, act, act_slope)).view_as(out_planes))


def wgrad_cl(row_srcs, col_srcs, row_segs, col_segs, B, T, *, scale=1.0, prec=0, splits=None, out=None):
    assert N.wgrad_supported(row_segs, col_segs), (row_segs, col_segs)
    rows = torch.cat([dec(row_srcs[si])[:, :, co:co + w] for si, co, w in row_segs], dim=2)          # [B,T,R]
    cols = torch.cat([_shift_rows(dec(col_srcs[si]), int(sh))[:, :, co:co + w] for si, sh, co, w in col_segs], dim=2)
    for t in list(row_srcs) + list(col_srcs):
        assert t.dim() == 4 and t.shape[0] == 2 and t.shape[1] == B and t.shape[2] == T and t.shape[3] % 8 == 0
    g = torch.einsum("btr,btc->rc", rows, cols) * scale
    g = g.to(torch.float32)
    if out is not None:
        out.copy_(g)
        return out
    return g


def lrelu_bwd(grad, act_planes, slope, *, addend=None, out_f32=None, out_planes=None, scale=1.0, prec=0):
    assert grad.numel() % 4 == 0
    a = dec(act_planes).reshape(grad.shape)
    v = torch.where(a > 0, grad.to(torch.float64), grad.to(torch.float64) * slope) * scale
    if addend is not None:
        v = v + addend.to(torch.float64)
    if out_f32 is not None:
        out_f32.copy_(v.to(torch.float32))
    if out_planes is not None:
        out_planes.copy_(enc(v).view_as(out_planes))


def colsum(planes, B, T, Nn, *, scale=1.0, prec=0):
    return (dec(planes).reshape(B * T, Nn).sum(0) * scale).to(torch.float32)


@contextlib.contextmanager
def emulated_native():
    names = ("split_nwc", "pack_weight", "mrf_finish", "conv_cl", "wgrad_cl", "lrelu_bwd", "colsum", "require_cuda")
    saved = {n: getattr(N, n) for n in names}
    try:
        N.split_nwc, N.pack_weight, N.mrf_finish, N.conv_cl = split_nwc, pack_weight, mrf_finish, conv_cl
        N.wgrad_cl, N.lrelu_bwd, N.colsum = wgrad_cl, lrelu_bwd, colsum
        N.require_cuda = lambda t, name="tensor": None
        yield
    finally:
        for n, f in saved.items():
            setattr(N, n, f)
