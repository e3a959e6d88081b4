"""Oracle: noise schedule, DDPM / PLMS / UniPC samplers, GaussianDiffusion.forward / train_step.
TEST INFRASTRUCTURE ONLY.  Reference: fish_diffusion/archs/diffsinger/diffusions/{diffusion,noise_predictor,uni_pc}.py.

Random numbers are never drawn here: every N(0,1) tensor the reference would draw is an explicit input,
so that both sides of a parity test consume identical noise.
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------ schedule
def get_noise_schedule_list(schedule_mode, timesteps, max_beta=0.01, s=0.008):
    """diffusion.py:18-31 (float64 numpy, exactly as the reference)."""
    if schedule_mode == "linear":
        return np.linspace(1e-4, max_beta, timesteps)
    if schedule_mode == "cosine":
        steps = timesteps + 1
        x = np.linspace(0, steps, steps)
        ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = 1 - (ac[1:] / ac[:-1])
        return np.clip(betas, a_min=0, a_max=0.999)
    raise NotImplementedError(schedule_mode)


def diffusion_tables(betas):
    """fp32 buffers of GaussianDiffusion (diffusion.py:70-88) and NaiveNoisePredictor (noise_predictor.py:28-71),
    computed in float64 and cast to float32 -- the bit-exact requirement of north_star applies to these."""
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "alphas_cumprod_prev": f32(ac_prev),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(post_var),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }


def sampling_chunks(num_timesteps, sampler_interval, skip_steps=0):
    """diffusion.py:234-240: arange(0, N-skip, interval).flip(0) as int64."""
    return np.arange(0, num_timesteps - skip_steps, sampler_interval, dtype=np.int64)[::-1].copy()


def norm_spec(x, spec_min, spec_max):
    """diffusion.py:315-316."""
    return (x - spec_min) / (spec_max - spec_min) * 2 - 1


def denorm_spec(x, spec_min, spec_max):
    """diffusion.py:318-319."""
    return (x + 1) / 2 * (spec_max - spec_min) + spec_min


# ------------------------------------------------------------------------------------------ DDPM
def naive_step(tab, x, t, eps, noise, clip_min=-1.0, clip_max=1.0):
    """NaiveNoisePredictor.forward (noise_predictor.py:73-104); `noise` replaces torch.randn_like."""
    dt = x.dtype
    g = lambda k: tab[k][t].astype(dt)
    x0 = g("sqrt_recip_alphas_cumprod") * x - g("sqrt_recipm1_alphas_cumprod") * eps
    x0 = np.clip(x0, clip_min, clip_max)
    mean = g("posterior_mean_coef1") * x0 + g("posterior_mean_coef2") * x
    logvar = g("posterior_log_variance_clipped")
    nonzero = 1.0 if t > 0 else 0.0
    return mean + nonzero * np.exp(0.5 * logvar) * noise


# ------------------------------------------------------------------------------------------ PLMS
def plms_x_pred(tab, x, noise_t, t, t_prev):
    """PLMSNoisePredictor.forward (noise_predictor.py:118-131)."""
    dt = x.dtype
    a_t = tab["alphas_cumprod"][t].astype(dt)
    a_prev = tab["alphas_cumprod"][t_prev].astype(dt)
    a_t_sq, a_prev_sq = np.sqrt(a_t), np.sqrt(a_prev)
    x_delta = (a_prev - a_t) * (
        (1 / (a_t_sq * (a_t_sq + a_prev_sq))) * x
        - 1 / (a_t_sq * (np.sqrt((1 - a_prev) * a_t) + np.sqrt((1 - a_t) * a_prev))) * noise_t
    )
    return x + x_delta


def plms_sample(tab, denoise, x, num_timesteps, sampler_interval, skip_steps=0, denoise_nomask=None):
    """diffusion.py:269-311 (PLMS branch of GaussianDiffusion.forward).  The stage-0 look-ahead call omits
    the masks (diffusion.py:285): `denoise_nomask`."""
    if denoise_nomask is None:
        denoise_nomask = denoise
    stage = 0
    noise_list = []
    for t in sampling_chunks(num_timesteps, sampler_interval, skip_steps):
        t = int(t)
        noise_pred = denoise(x, np.array([t], dtype=np.int64))
        t_prev = t - sampler_interval
        t_prev = t_prev * (t_prev > 0)
        if stage == 0:
            x_pred = plms_x_pred(tab, x, noise_pred, t, t_prev)
            noise_pred_prev = denoise_nomask(x_pred, np.array([t_prev], dtype=np.int64))
            prime = (noise_pred + noise_pred_prev) / 2
        elif stage == 1:
            prime = (noise_pred * 3 - noise_list[-1]) / 2
        elif stage == 2:
            prime = (noise_pred * 23 - noise_list[-1] * 16 + noise_list[-2] * 5) / 12
        else:
            prime = (noise_pred * 55 - noise_list[-1] * 59 + noise_list[-2] * 37 - noise_list[-3] * 9) / 24
        if stage < 3:
            noise_list.append(noise_pred)
            stage += 1
        else:
            noise_list = noise_list[-2:] + [noise_pred]
        x = plms_x_pred(tab, x, prime, t, t_prev)
    return x


# ------------------------------------------------------------------------------------------ UniPC
def torch_linspace_f32(start, end, steps):
    """torch.linspace(..., dtype=float32) bit-for-bit (ATen RangeFactories linspace kernel): step in fp32, first
    half start + step*i, second half end - step*(steps-1-i), each evaluated with one fused multiply-add."""
    start, end = np.float32(start), np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    lo = (np.float64(start) + np.float64(step) * i).astype(np.float32)
    hi = (np.float64(end) - np.float64(step) * (steps - i - 1)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def interpolate_fn(x, xp, yp):
    """uni_pc.py:826-875 for C == 1: piecewise-linear through (xp, yp), outermost segments extrapolate.
    x scalar, xp/yp 1-D ascending."""
    K = xp.shape[0]
    x_idx = int(np.searchsorted(xp, x, side="left"))
    if x_idx == 0:
        i0 = 0
    elif x_idx == K:
        i0 = K - 2
    else:
        i0 = x_idx - 1
    sx, ex, sy, ey = xp[i0], xp[i0 + 1], yp[i0], yp[i0 + 1]
    return sy + (x - sx) * (ey - sy) / (ex - sx)


class NoiseScheduleVP:
    """uni_pc.py:6-197, schedule='discrete' only (what UNIPCNoisePredictor builds, noise_predictor.py:155-158)."""

    def __init__(self, betas, dtype=np.float32):
        log_alphas = 0.5 * np.cumsum(np.log(1 - np.asarray(betas, dtype=np.float64)))
        self.total_N = len(log_alphas)
        self.T = 1.0
        self.dtype = dtype
        self.t_array = torch_linspace_f32(0.0, 1.0, self.total_N + 1)[1:].astype(dtype)
        self.log_alpha_array = log_alphas.astype(np.float32).astype(dtype)

    def marginal_log_mean_coeff(self, t):
        return interpolate_fn(self.dtype(t), self.t_array, self.log_alpha_array)

    def marginal_alpha(self, t):
        return np.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return np.sqrt(1.0 - np.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lmc = self.marginal_log_mean_coeff(t)
        return lmc - 0.5 * np.log(1.0 - np.exp(2.0 * lmc))


def unipc_sample(betas, denoise, x, sampler_interval=10, order=2, dtype=np.float32):
    """UNIPCNoisePredictor.forward (noise_predictor.py:176-222) -> UniPC.sample (uni_pc.py:703-818) with
    variant='bh2', data prediction, method='multistep', skip_type='time_uniform', lower_order_final=True,
    via multistep_uni_pc_bh_update (uni_pc.py:583-701).  `denoise(x, t_input[B float])` returns eps.
    Scalar coefficient math runs in `dtype` (float32 mirrors the reference's torch scalars)."""
    ns = NoiseScheduleVP(betas, dtype=dtype)
    steps = ns.total_N // sampler_interval
    N = ns.total_N
    f = dtype
    B = x.shape[0]

    def model_fn(xx, t):
        # model_wrapper.noise_pred_fn (uni_pc.py:214-240) + UniPC.data_prediction_fn (uni_pc.py:327-339)
        t_input = (f(t) - f(1.0 / N)) * f(N)
        eps = denoise(xx, np.full((B,), t_input, dtype=f))
        alpha_t, sigma_t = ns.marginal_alpha(t), ns.marginal_std(t)
        return (xx - sigma_t * eps) / alpha_t

    def bh_update(xx, model_prev_list, t_prev_list, t, upd_order, use_corrector):
        t_prev_0 = t_prev_list[-1]
        lambda_prev_0, lambda_t = ns.marginal_lambda(t_prev_0), ns.marginal_lambda(t)
        model_prev_0 = model_prev_list[-1]
        sigma_prev_0, sigma_t = ns.marginal_std(t_prev_0), ns.marginal_std(t)
        alpha_t = np.exp(ns.marginal_log_mean_coeff(t))
        h = lambda_t - lambda_prev_0
        rks, D1s = [], []
        for i in range(1, upd_order):
            t_prev_i = t_prev_list[-(i + 1)]
            model_prev_i = model_prev_list[-(i + 1)]
            rk = (ns.marginal_lambda(t_prev_i) - lambda_prev_0) / h
            rks.append(rk)
            D1s.append((model_prev_i - model_prev_0) / rk)
        rks.append(f(1.0))
        rks = np.array(rks, dtype=f)
        hh = -h
        h_phi_1 = np.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        B_h = np.expm1(hh)
        R, b = [], []
        for i in range(1, upd_order + 1):
            R.append(np.power(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        R = np.stack(R).astype(f)
        b = np.array(b, dtype=f)
        use_predictor = len(D1s) > 0
        rhos_p = np.array([0.5], dtype=f) if upd_order == 2 else None
        if use_corrector:
            rhos_c = np.array([0.5], dtype=f) if upd_order == 1 else np.linalg.solve(R, b).astype(f)
        x_t_ = sigma_t / sigma_prev_0 * xx - alpha_t * h_phi_1 * model_prev_0
        pred_res = sum(rhos_p[k] * D1s[k] for k in range(len(D1s))) if use_predictor else 0
        x_t = x_t_ - alpha_t * B_h * pred_res
        model_t = None
        if use_corrector:
            model_t = model_fn(x_t, t)
            corr_res = sum(rhos_c[k] * D1s[k] for k in range(len(D1s))) if D1s else 0
            D1_t = model_t - model_prev_0
            x_t = x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)
        return x_t, model_t

    t_0, t_T = 1.0 / N, ns.T
    timesteps = torch_linspace_f32(t_T, t_0, steps + 1).astype(f)
    t = timesteps[0]
    t_prev_list = [t]
    model_prev_list = [model_fn(x, t)]
    for step in range(1, order):
        t = timesteps[step]
        x, model_x = bh_update(x, model_prev_list, t_prev_list, t, step, True)
        if model_x is None:
            model_x = model_fn(x, t)
        t_prev_list.append(t)
        model_prev_list.append(model_x)
    for step in range(order, steps + 1):
        t = timesteps[step]
        step_order = min(order, steps + 1 - step)
        use_corrector = step != steps
        x, model_x = bh_update(x, model_prev_list, t_prev_list, t, step_order, use_corrector)
        for i in range(order - 1):
            t_prev_list[i] = t_prev_list[i + 1]
            model_prev_list[i] = model_prev_list[i + 1]
        t_prev_list[-1] = t
        if step < steps:
            if model_x is None:
                model_x = model_fn(x, t)
            model_prev_list[-1] = model_x
    return x


# ------------------------------------------------------------------------------------------ GaussianDiffusion
def diffusion_forward(betas, denoise_bmt, features, x_T, step_noises=None, sampler_interval=10, skip_steps=0,
                      noise_predictor="naive", spec_min=-5.0, spec_max=0.0, x_masks=None, cond_masks=None,
                      dtype=np.float64):
    """GaussianDiffusion.forward (diffusion.py:196-313).  features [B,T,E]; x_T [B,M,T] is the initial noise
    (diffusion.py:219-222); step_noises[i] is the N(0,1) draw of the i-th naive step (noise_predictor.py:101).
    denoise_bmt(x[B,M,T], t, cond[B,E,T], x_masks, cond_masks) -> eps[B,M,T].  Returns mel [B,T,M]."""
    betas = np.asarray(betas, dtype=np.float64)
    tab = diffusion_tables(betas)
    N = len(betas)
    cond = np.transpose(np.asarray(features, dtype=dtype), (0, 2, 1))
    x = np.asarray(x_T, dtype=dtype)
    den = lambda xx, t: denoise_bmt(xx, t, cond, x_masks, cond_masks)
    if noise_predictor == "naive":
        for i, t in enumerate(sampling_chunks(N, sampler_interval, skip_steps)):
            t = int(t)
            eps = den(x, np.array([t], dtype=np.int64))
            x = naive_step(tab, x, t, eps, np.asarray(step_noises[i], dtype=dtype))
    elif noise_predictor == "unipc":
        x = unipc_sample(betas, den, x, sampler_interval=sampler_interval, dtype=np.float32 if dtype == np.float32 else np.float64)
    elif noise_predictor == "plms":
        den_nomask = lambda xx, t: denoise_bmt(xx, t, cond, None, None)
        x = plms_sample(tab, den, x, N, sampler_interval, skip_steps, denoise_nomask=den_nomask)
    else:
        raise NotImplementedError(noise_predictor)
    return denorm_spec(np.transpose(x, (0, 2, 1)), spec_min, spec_max)


def q_sample(tab, x_start, t, noise):
    """diffusion.py:120-127; t int64 [B]."""
    dt = x_start.dtype
    a = tab["sqrt_alphas_cumprod"][t].astype(dt).reshape(-1, 1, 1)
    s = tab["sqrt_one_minus_alphas_cumprod"][t].astype(dt).reshape(-1, 1, 1)
    return a * x_start + s * noise


def mel_loss(kind, noise, epsilon):
    """diffusion.py:153-170."""
    d = noise - epsilon
    if kind == "l1":
        return np.mean(np.abs(d))
    if kind == "l2":
        return np.mean(d * d)
    if kind == "smoothed-l1":
        ad = np.abs(d)
        return np.mean(np.where(ad < 1.0, 0.5 * d * d, ad - 0.5))
    raise NotImplementedError(kind)
