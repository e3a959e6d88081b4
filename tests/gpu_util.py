"""Helpers for the -m gpu parity tests."""
import numpy as np
import torch

from fish_diffusion_b200 import _native as N


def planes_to_f64(planes: torch.Tensor, prec: int) -> np.ndarray:
    """split planes int16 [2, ...] (device) -> exact float64 value hi + lo."""
    p = planes.cpu()
    if prec == N.PREC_F16:
        hi = p[0].view(torch.float16).to(torch.float64)
        lo = p[1].view(torch.float16).to(torch.float64)
    else:
        hi = p[0].view(torch.bfloat16).to(torch.float64)
        lo = p[1].view(torch.bfloat16).to(torch.float64)
    return (hi + lo).numpy()


def tap_gemm_ref(a, w, shifts, bias=None):
    """a [B,T,Ci] float64, w [N, ntaps*Ci] float64 -> [B,T,N]: sum_j a[b, t+shift_j, :] @ w[:, j*Ci:(j+1)*Ci].T"""
    B, T, Ci = a.shape
    out = np.zeros((B, T, w.shape[0]))
    for j, s in enumerate(shifts):
        seg = np.zeros_like(a)
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            seg[:, lo:hi] = a[:, lo + s:hi + s]
        out += seg @ w[:, j * Ci:(j + 1) * Ci].T
    if bias is not None:
        out += bias
    return out


def dev():
    return torch.device("cuda:0")
