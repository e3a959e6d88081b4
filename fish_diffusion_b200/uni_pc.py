"""UniPC (order-2 multistep, B(h)=expm1 'bh2', data prediction) host logic for the native sampler.

Replaces the tensor updates of the reference's ``UniPC.sample`` / ``multistep_uni_pc_bh_update``
(reference fish_diffusion/archs/diffsinger/diffusions/uni_pc.py:583-818, driven by
noise_predictor.py:176-222 with steps = N // sampler_interval, order=2, skip_type='time_uniform',
lower_order_final=True).  The scalar coefficient math is float32 like the reference's torch scalars and stays on the
host; each predictor / corrector update is ONE fused fd_lincomb kernel over the channels-last state instead of the
reference's chain of broadcasted elementwise ops and einsums.
"""
from __future__ import annotations

import numpy as np

from . import _native as N

F = np.float32


def linspace_f32(start, end, steps):
    """torch.linspace(start, end, steps) in float32, bit for bit (two-sided fill, one fma per element)."""
    start, end = F(start), F(end)
    step = F((end - start) / F(steps - 1))
    i = np.arange(steps)
    lo = (np.float64(start) + np.float64(step) * i).astype(np.float32)
    hi = (np.float64(end) - np.float64(step) * (steps - i - 1)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


class NoiseScheduleVP:
    """Discrete VP schedule of the reference (uni_pc.py:6-197, schedule='discrete'): float32 key points
    t_i = (i+1)/N and log_alpha_i = 0.5*cumsum(log(1-beta)) (float64 -> float32), piecewise-linear in between."""

    def __init__(self, betas):
        log_alphas = 0.5 * np.cumsum(np.log(1 - np.asarray(betas, dtype=np.float64)))
        self.schedule = "discrete"
        self.total_N = len(log_alphas)
        self.T = 1.0
        self.t_array = linspace_f32(0.0, 1.0, self.total_N + 1)[1:]
        self.log_alpha_array = log_alphas.astype(np.float32)

    def marginal_log_mean_coeff(self, t):
        """interpolate_fn (uni_pc.py:826-875) specialised to one query: the outermost segments extrapolate."""
        xp, yp = self.t_array, self.log_alpha_array
        t = F(t)
        K = xp.shape[0]
        idx = int(np.searchsorted(xp, t, side="left"))
        i0 = 0 if idx == 0 else (K - 2 if idx == K else idx - 1)
        return F(yp[i0] + (t - xp[i0]) * (yp[i0 + 1] - yp[i0]) / (xp[i0 + 1] - xp[i0]))

    def marginal_alpha(self, t):
        return F(np.exp(self.marginal_log_mean_coeff(t)))

    def marginal_std(self, t):
        return F(np.sqrt(F(1.0) - np.exp(F(2.0) * self.marginal_log_mean_coeff(t))))

    def marginal_lambda(self, t):
        lmc = self.marginal_log_mean_coeff(t)
        return F(lmc - F(0.5) * np.log(F(1.0) - np.exp(F(2.0) * lmc)))


def _update_coefs(ns: NoiseScheduleVP, t_prev_list, t, order, use_corrector):
    """Scalar part of multistep_uni_pc_bh_update (uni_pc.py:583-668) for variant bh2 / predict_x0.
    Returns (A, base, pred, corr) with
        x_t_   = A*x + base*m0
        x_pred = x_t_ + pred[0]*m0 + pred[1]*m1
        x_corr = x_t_ + corr[0]*m0 + corr[1]*m1 + corr[2]*m_t
    (m0 = newest stored model output, m1 the one before; unused entries are 0)."""
    t_prev_0 = t_prev_list[-1]
    lambda_prev_0, lambda_t = ns.marginal_lambda(t_prev_0), ns.marginal_lambda(t)
    sigma_prev_0, sigma_t = ns.marginal_std(t_prev_0), ns.marginal_std(t)
    alpha_t = F(np.exp(ns.marginal_log_mean_coeff(t)))
    h = F(lambda_t - lambda_prev_0)
    rks = []
    for i in range(1, order):
        rks.append(F((ns.marginal_lambda(t_prev_list[-(i + 1)]) - lambda_prev_0) / h))
    rks.append(F(1.0))
    rks = np.array(rks, dtype=np.float32)
    hh = F(-h)
    h_phi_1 = F(np.expm1(hh))
    h_phi_k = F(h_phi_1 / hh - F(1))
    B_h = F(np.expm1(hh))
    factorial_i = 1
    R, b = [], []
    for i in range(1, order + 1):
        R.append(np.power(rks, i - 1).astype(np.float32))
        b.append(F(h_phi_k * F(factorial_i) / B_h))
        factorial_i *= i + 1
        h_phi_k = F(h_phi_k / hh - F(1.0 / factorial_i))
    R = np.stack(R).astype(np.float32)
    b = np.array(b, dtype=np.float32)
    A = F(sigma_t / sigma_prev_0)
    base = F(-alpha_t * h_phi_1)
    g = F(alpha_t * B_h)
    pred = [F(0), F(0)]
    corr = [F(0), F(0), F(0)]
    if order == 2:
        rk = rks[0]
        pred = [F(g * F(0.5) / rk), F(-g * F(0.5) / rk)]        # rhos_p = 0.5 (uni_pc.py:641-643)
    if use_corrector:
        if order == 1:
            rc = np.array([0.5], dtype=np.float32)                # uni_pc.py:652-653
            corr = [F(g * rc[0]), F(0), F(-g * rc[0])]
        else:
            rc = np.linalg.solve(R, b).astype(np.float32)
            rk = rks[0]
            corr = [F(g * (rc[0] / rk + rc[1])), F(-g * rc[0] / rk), F(-g * rc[1])]
    return A, base, pred, corr


def unipc_sample_native(ns: NoiseScheduleVP, x, x_planes, denoise, sampler_interval, prec, order=2, progress=False):
    """x fp32 [B,T,M] channels-last (updated and returned), x_planes its split planes;
    denoise(planes, t_input: float, out) writes eps into `out`."""
    import torch
    from .diffusion import lincomb

    steps = ns.total_N // sampler_interval
    NN = ns.total_N
    ts = linspace_f32(ns.T, 1.0 / NN, steps + 1)
    bar = None
    if progress:
        from tqdm import tqdm
        bar = tqdm(total=steps)

    eps = torch.empty_like(x)
    x_pred = torch.empty_like(x)
    xp_planes = torch.empty_like(x_planes)
    bufs = [torch.empty_like(x) for _ in range(3)]   # rotating storage for model outputs

    def model_fn(src, src_planes, t, out):
        # model_wrapper.noise_pred_fn + UniPC.data_prediction_fn (uni_pc.py:214-240, 327-339)
        t_input = F((F(t) - F(1.0 / NN)) * F(NN))
        denoise(src_planes, float(t_input), eps)
        alpha_t, sigma_t = ns.marginal_alpha(t), ns.marginal_std(t)
        lincomb(out, [(F(1) / alpha_t, src), (-sigma_t / alpha_t, eps)])
        if bar is not None:
            bar.update(1)
        return out

    t_prev_list = [ts[0]]
    m0 = model_fn(x, x_planes, ts[0], bufs[0])
    m1 = None
    free = [bufs[1], bufs[2]]
    for step in range(1, steps + 1):
        t = ts[step]
        if step < order:
            upd_order, use_corrector = step, True
        else:
            upd_order, use_corrector = min(order, steps + 1 - step), step != steps
        A, base, pred, corr = _update_coefs(ns, t_prev_list, t, upd_order, use_corrector)
        if use_corrector:
            terms = [(A, x), (base + pred[0], m0)]
            if upd_order == 2:
                terms.append((pred[1], m1))
            lincomb(x_pred, terms, planes=xp_planes, prec=prec)
            mt = model_fn(x_pred, xp_planes, t, free.pop())
            terms = [(A, x), (base + corr[0], m0)]
            if upd_order == 2:
                terms.append((corr[1], m1))
            terms.append((corr[2], mt))
            lincomb(x, terms)
            if m1 is not None:
                free.append(m1)
            m1, m0 = m0, mt
        else:
            terms = [(A, x), (base + pred[0], m0)]
            if upd_order == 2:
                terms.append((pred[1], m1))
            lincomb(x, terms)
        if len(t_prev_list) < order:
            t_prev_list.append(t)
        else:
            t_prev_list = t_prev_list[1:] + [t]
    if bar is not None:
        bar.close()
    return x
