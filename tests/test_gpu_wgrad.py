"""GPU parity of the direct weight-gradient kernel (fd_wgrad_cl: tcgen05 with MN-major operands read straight from
channels-last planes) against float64 on the exact plane values."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import _native as N
from gpu_util import dev, planes_to_f64

pytestmark = pytest.mark.gpu


def hi_to_f64(planes, pc):
    p = planes[0].cpu()
    return p.view(torch.float16 if pc == N.PREC_F16 else torch.bfloat16).to(torch.float64).numpy()


def shifted(a, s):
    """a [B,T,C] -> a[b, t+s, c] with zeros outside [0,T)"""
    B, T, C = a.shape
    out = np.zeros_like(a)
    lo, hi = max(0, -s), min(T, T - s)
    if hi > lo:
        out[:, lo:hi] = a[:, lo + s:hi + s]
    return out


CASES = [
    # B, T, row channel counts, row segs (src, coff, width), col channel counts, col segs (src, shift, coff, width)
    (3, 200, [128], [(0, 0, 128)], [64, 64], [(0, -2, 0, 64), (0, 0, 0, 64), (0, 2, 0, 64), (1, 0, 0, 64)]),
    (2, 77, [64], [(0, 0, 64)], [128], [(0, 0, 0, 128)]),
    (4, 130, [128, 128], [(0, 0, 128), (1, 0, 128)], [192], [(0, 0, 0, 192)]),
    (2, 1000, [256], [(0, 0, 256)], [128, 256], [(0, -64, 0, 128), (0, 0, 0, 128), (0, 64, 0, 128), (1, 0, 0, 256)]),
    (5, 64, [192], [(0, 64, 128)], [320], [(0, 3, 64, 256)]),
]


@pytest.mark.parametrize("prec", ["f16", "bf16", "f16x1"])
@pytest.mark.parametrize("splits", [None, 1])
@pytest.mark.parametrize("case", CASES)
def test_wgrad_direct_vs_float64(case, prec, splits):
    B, T, rowC, row_segs, colC, col_segs = case
    pc, mma = N.prec_code(prec), N.mma_code(prec)
    rng = np.random.RandomState(B * 1000 + T)
    rows = [N.split_nwc(torch.from_numpy(rng.randn(B, T, c).astype(np.float32)).to(dev()), pc) for c in rowC]
    cols = [N.split_nwc(torch.from_numpy(rng.randn(B, T, c).astype(np.float32)).to(dev()), pc) for c in colC]
    out = N.wgrad_cl(rows, cols, row_segs, col_segs, B, T, scale=0.25, prec=mma, splits=splits)
    torch.cuda.synchronize()
    conv = hi_to_f64 if prec.endswith("x1") else planes_to_f64
    r64 = [conv(p, pc) for p in rows]
    c64 = [conv(p, pc) for p in cols]
    Rm = np.concatenate([r64[s][:, :, co:co + w] for s, co, w in row_segs], axis=2)
    Cm = np.concatenate([shifted(c64[s], sh)[:, :, co:co + w] for s, sh, co, w in col_segs], axis=2)
    ref = 0.25 * np.einsum("btr,btc->rc", Rm, Cm)
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    tol = 2e-5 if pc == N.PREC_F16 else 5e-5   # fp32 TMEM accumulation over K = B*T up to 2000 terms
    assert rel_l2(got, ref) < tol, (rel_l2(got, ref), np.abs(got - ref).max())


def test_colsum_edges():
    B, T, Nn, e = 3, 100, 192, 17
    rng = np.random.RandomState(4)
    a = torch.from_numpy(rng.randn(B, T, Nn).astype(np.float32)).to(dev())
    pl = N.split_nwc(a, N.PREC_F16)
    out = torch.zeros((2, B, Nn), dtype=torch.float32, device=dev())
    N.check(N.lib().fd_colsum_edges(N.ptr(pl), N.ptr(out), B, T, Nn, e, 0.5, N.PREC_F16, N.stream_ptr(dev())),
            "fd_colsum_edges")
    v = planes_to_f64(pl, N.PREC_F16)
    assert rel_l2(out[0].cpu().numpy(), 0.5 * v[:, :e].sum(1)) < 1e-6
    assert rel_l2(out[1].cpu().numpy(), 0.5 * v[:, T - e:].sum(1)) < 1e-6
