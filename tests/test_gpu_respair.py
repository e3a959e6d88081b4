"""GPU parity: the fused ResBlock1 pair kernel (fd_respair_fwd, csrc/fd_respair_tc.cu) through the C ABI against a
float64 restatement of  x' = x + c2(lrelu(c1(lrelu(x))))  (reference models.py:103-110) on the exact operand values
read back from the split planes.  Covers every channel width, the (k, dilation) pairs of config_v1, ragged T (tiles
that end inside an item, items shorter than one tile), B > 1 (no leakage across items), all three output modes."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import _native as N
from gpu_util import dev, planes_to_f64, tap_gemm_ref

pytestmark = pytest.mark.gpu

SLOPE = 0.1


def lrelu(x, s=SLOPE):
    return np.where(x >= 0, x, x * s)


def pair_ref(p_in, w1, b1, w2, b2, k1, d1, k2):
    """p_in = lrelu(x) [B,T,C] float64 (exact plane values); conv weights [C][k*C] tap-major."""
    x = np.where(p_in >= 0, p_in, p_in / SLOPE)
    s1 = [(j - (k1 - 1) // 2) * d1 for j in range(k1)]
    s2 = [(j - (k2 - 1) // 2) for j in range(k2)]
    mid = lrelu(tap_gemm_ref(p_in, w1, s1, b1))
    return x + tap_gemm_ref(mid, w2, s2, b2)


def make_case(B, T, C, k, d, seed):
    rng = np.random.RandomState(seed)
    x = rng.randn(B, T, C).astype(np.float32) * 1.5
    w1 = (rng.randn(C, k * C) / np.sqrt(k * C)).astype(np.float32)
    w2 = (rng.randn(C, k * C) / np.sqrt(k * C)).astype(np.float32)
    b1 = (rng.randn(C) * 0.3).astype(np.float32)
    b2 = (rng.randn(C) * 0.3).astype(np.float32)
    return x, w1, b1, w2, b2


def run_pair(x, w1, b1, w2, b2, k, d, prec="f16", out_slope=SLOPE, planes_scale=1.0, kmask1=0, kmask2=0):
    pc = N.prec_code(prec)
    B, T, C = x.shape
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    xin = t(x)
    pa = N.split_nwc(torch.where(xin >= 0, xin, xin * SLOPE), pc)
    s1, s2 = N.pow2_scale(t(w1)), N.pow2_scale(t(w2))
    w1p, w2p = N.pack_weight(t(w1), pc, s1), N.pack_weight(t(w2), pc, s2)
    out = torch.zeros((2, B, T, C), dtype=torch.int16, device=dev())
    N.respair(pa, w1p, w2p, t(b1), t(b2), B, T, C, k, d, k, out_planes=out, w1_inv_scale=1.0 / s1,
              w2_inv_scale=1.0 / s2, in_slope=SLOPE, out_slope=out_slope, planes_scale=planes_scale,
              prec=N.mma_code(prec), kmask1=kmask1, kmask2=kmask2)
    torch.cuda.synchronize()
    ref = pair_ref(planes_to_f64(pa, pc), planes_to_f64(w1p, pc) / s1, b1.astype(np.float64),
                   planes_to_f64(w2p, pc) / s2, b2.astype(np.float64), k, d, k)
    return out, ref, pc


KD = [(3, 1), (3, 3), (3, 5), (7, 1), (7, 3), (7, 5), (11, 1), (11, 3), (11, 5)]


@pytest.mark.parametrize("C", [16, 32, 64, 128])
@pytest.mark.parametrize("kd", KD)
def test_pair_planes_out(C, kd):
    k, d = kd
    assert N.respair_supported(C, k, d, k)
    B, T = 2, 300 + 7 * k + d          # tiles end inside the item; the second item starts on a fresh tile
    x, w1, b1, w2, b2 = make_case(B, T, C, k, d, 100 * C + 10 * k + d)
    out, ref, pc = run_pair(x, w1, b1, w2, b2, k, d)
    got = planes_to_f64(out, pc)
    want = lrelu(ref)
    # mid activations are re-split to 22-bit planes inside the kernel: 2^-22 relative per element, summed over k*C
    assert rel_l2(got, want) < 1e-5, (rel_l2(got, want), np.abs(got - want).max())   # measured <= 4e-6 (K up to 1408, two chained convs)


def test_mrf_finish():
    """fd_mrf_finish: lrelu(mean_i invlrelu(p_i)) on plane tensors (models.py:426-434)."""
    rng = np.random.RandomState(2)
    xs = [torch.from_numpy(rng.randn(3, 50, 16).astype(np.float32)).to(dev()) for _ in range(3)]
    ps = [N.split_nwc(torch.where(x >= 0, x, x * SLOPE), N.PREC_F16) for x in xs]
    out = torch.empty_like(ps[0])
    N.mrf_finish(ps, out, in_slope=SLOPE, scale=1.0 / 3, out_slope=0.01)
    torch.cuda.synchronize()
    vals = [planes_to_f64(p, N.PREC_F16) for p in ps]
    want = lrelu(sum(np.where(v >= 0, v, v / SLOPE) for v in vals) / 3, 0.01)
    assert rel_l2(planes_to_f64(out, N.PREC_F16), want) < 1e-6


@pytest.mark.parametrize("T", [1, 5, 245, 246, 247, 256, 1000])
def test_pair_ragged_lengths(T):
    C, k, d = 64, 11, 5                 # r_out = 2*128 - 10 = 246: items shorter than / equal to / just above one tile
    x, w1, b1, w2, b2 = make_case(2, T, C, k, d, T)
    out, ref, pc = run_pair(x, w1, b1, w2, b2, k, d)
    assert rel_l2(planes_to_f64(out, pc), lrelu(ref)) < 1e-5


def test_pair_items_are_independent():
    """Item 1 of a batch equals the same item run alone (halo rows never cross an item boundary)."""
    C, k, d = 32, 11, 5
    x, w1, b1, w2, b2 = make_case(3, 200, C, k, d, 9)
    full, _, pc = run_pair(x, w1, b1, w2, b2, k, d)
    solo, _, _ = run_pair(x[1:2], w1, b1, w2, b2, k, d)
    assert torch.equal(full[:, 1], solo[:, 0])


@pytest.mark.parametrize("prec", ["bf16", "f16x1"])
def test_pair_other_precisions(prec):
    C, k, d = 128, 7, 3
    x, w1, b1, w2, b2 = make_case(2, 300, C, k, d, 11)
    out, ref, pc = run_pair(x, w1, b1, w2, b2, k, d, prec=prec, out_slope=0.01, planes_scale=1.0 / 3)
    want = lrelu(ref, 0.01) / 3
    tol = 5e-5 if prec == "bf16" else 2e-3     # single product: 11-bit operands
    assert rel_l2(planes_to_f64(out, pc), want) < tol


def test_pair_large_batch_many_tiles():
    """More tiles than SMs (persistent loop, ring phases wrap many times) at the narrowest width."""
    C, k, d = 16, 3, 1
    x, w1, b1, w2, b2 = make_case(4, 40000, C, k, d, 21)
    out, ref, pc = run_pair(x, w1, b1, w2, b2, k, d)
    assert rel_l2(planes_to_f64(out, pc), lrelu(ref)) < 1e-5


def block_mask(w, C, k):
    """fd_respair_desc.kmask: bit tap*(C/16)+s <=> input channels [16s, 16s+16) of that tap hold a non-zero weight."""
    nz = (w.reshape(C, k, C // 16, 16) != 0).any(axis=(0, 3)).reshape(-1)
    return sum(1 << i for i, b in enumerate(nz) if b)


@pytest.mark.parametrize("C,k,d", [(32, 11, 1), (32, 7, 3), (32, 27, 1), (64, 7, 1), (128, 3, 1)])
def test_pair_block_sparse_hint(C, k, d):
    """Block-sparse weights (the shape of the time-folded C = 16 kernels: most (tap, 16-channel) blocks are zero, including
    the first one) with the sparsity hint: same planes as the dense run of the same weights, and the float64 reference."""
    if not N.respair_supported(C, k, d, min(k, 17)):
        pytest.skip("shape outside the kernel's range")
    k2 = min(k, 7)
    rng = np.random.RandomState(5 * C + k)
    x, w1, b1, _, b2 = make_case(2, 700, C, k, d, 77)
    w2 = (rng.randn(C, k2 * C) / np.sqrt(k2 * C)).astype(np.float32)
    keep1 = rng.rand(k, C // 16) < 0.4
    keep1[0, 0] = False                       # the first K16 step of the first tap is skipped: accumulate flag logic
    keep1[k // 2, :] = True
    keep2 = rng.rand(k2, C // 16) < 0.6
    keep2[k2 // 2, 0] = True
    w1 = (w1.reshape(C, k, C // 16, 16) * keep1[None, :, :, None]).reshape(C, k * C).astype(np.float32)
    w2 = (w2.reshape(C, k2, C // 16, 16) * keep2[None, :, :, None]).reshape(C, k2 * C).astype(np.float32)
    pc = N.prec_code("f16")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    xin = t(x)
    pa = N.split_nwc(torch.where(xin >= 0, xin, xin * SLOPE), pc)
    s1, s2 = N.pow2_scale(t(w1)), N.pow2_scale(t(w2))
    w1p, w2p = N.pack_weight(t(w1), pc, s1), N.pack_weight(t(w2), pc, s2)
    outs = []
    for m1, m2 in ((0, 0), (block_mask(w1, C, k), block_mask(w2, C, k2))):
        out = torch.zeros((2,) + x.shape, dtype=torch.int16, device=dev())
        N.respair(pa, w1p, w2p, t(b1), t(b2), x.shape[0], x.shape[1], C, k, d, k2, out_planes=out, w1_inv_scale=1.0 / s1,
                  w2_inv_scale=1.0 / s2, in_slope=SLOPE, out_slope=SLOPE, prec=N.mma_code("f16"), kmask1=m1, kmask2=m2)
        outs.append(out)
    torch.cuda.synchronize()
    ref = pair_ref(planes_to_f64(pa, pc), planes_to_f64(w1p, pc) / s1, b1.astype(np.float64),
                   planes_to_f64(w2p, pc) / s2, b2.astype(np.float64), k, d, k2)
    assert rel_l2(planes_to_f64(outs[1], pc), lrelu(ref)) < 1e-5
    # skipping zero blocks only removes exact zeros from fp32 sums
    assert rel_l2(planes_to_f64(outs[1], pc), planes_to_f64(outs[0], pc)) < 1e-7
