"""Oracle: NSF-HiFiGAN generator (reference fish_diffusion/modules/vocoders/nsf_hifigan/models.py:1-448).
TEST INFRASTRUCTURE ONLY.  Random draws (rand_ini, SineGen noise) are explicit inputs."""
import math

import numpy as np

from .wavenet import conv1d

LRELU_SLOPE = 0.1


def get_padding(kernel_size, dilation=1):
    """models.py:23-24."""
    return int((kernel_size * dilation - dilation) / 2)


def lrelu(x, slope):
    return np.where(x > 0, x, x * slope)


def fold_weight_norm(g, v):
    """torch weight_norm (dim=0): w = g * v / ||v|| with the norm over all dims but 0 (models.py:440-448 removes it)."""
    n = np.sqrt(np.sum(v.astype(np.float64) ** 2, axis=tuple(range(1, v.ndim)), keepdims=True))
    return (g.astype(np.float64) * v.astype(np.float64) / n).astype(v.dtype)


def conv_transpose1d(x, w, b, stride, padding):
    """torch ConvTranspose1d: x [B,Ci,T], w [Ci,Co,K]."""
    B, Ci, T = x.shape
    _, Co, K = w.shape
    full = np.zeros((B, Co, (T - 1) * stride + K), dtype=x.dtype)
    for k in range(K):
        full[:, :, k:k + (T - 1) * stride + 1:stride] += np.einsum("co,bct->bot", w[:, :, k], x, optimize=True)
    To = (T - 1) * stride - 2 * padding + K
    out = full[:, :, padding:padding + To]
    return out + b[None, :, None]


def f0_upsample(f0, hop):
    """F.interpolate(f0[:,None], size=T*hop, mode='linear') (models.py:411-413), fp32 like ATen
    (area_pixel_compute_source_index, align_corners=False).  f0 [B,T] -> [B,T*hop] float32."""
    f0 = np.asarray(f0, dtype=np.float32)
    B, T = f0.shape
    S = T * hop
    scale = np.float32(T) / np.float32(S)
    s = np.arange(S, dtype=np.float32)
    src = scale * (s + np.float32(0.5)) - np.float32(0.5)
    src = np.maximum(src, np.float32(0))
    i0 = np.minimum(src.astype(np.int64), T - 1)
    i1 = i0 + (i0 < T - 1)
    lam = np.clip(src - i0.astype(np.float32), np.float32(0), np.float32(1))
    w0 = np.float32(1) - lam
    # ATen evaluates w0*x0 + lam*x1 as fma(w0, x0, round(lam*x1)) -- verified bit-for-bit against torch 2.11 CPU
    t1 = (lam[None] * f0[:, i1]).astype(np.float32)
    return (t1.astype(np.float64) + w0[None].astype(np.float64) * f0[:, i0].astype(np.float64)).astype(np.float32)


def sinegen(f0_up, sampling_rate, rand_ini, noise, harmonic_num=8, sine_amp=0.1, noise_std=0.003, mode="exact"):
    """SineGen.forward / _f02sine (models.py:201-294).  f0_up [B,S] fp32, rand_ini [B,H] (col 0 must be 0),
    noise [B,S,H] ~ N(0,1).  Returns sine_waves [B,S,H] float64.

    mode="ref32": the reference's float32 recipe verbatim (two fp32 cumsums + wrap detection).
    mode="exact": the per-sample increments rad are the reference's fp32 values, but the phase is their exact
                  running sum mod 1 (what the fp32 recipe approximates; the CUDA kernel computes this)."""
    H = harmonic_num + 1
    f0_up = np.asarray(f0_up, dtype=np.float32)
    B, S = f0_up.shape
    f0_buf = np.zeros((B, S, H), dtype=np.float32)
    f0_buf[:, :, 0] = f0_up
    for idx in range(harmonic_num):
        f0_buf[:, :, idx + 1] = f0_buf[:, :, 0] * np.float32(idx + 2)
    rad = np.mod(f0_buf / np.float32(sampling_rate), np.float32(1))
    rad[:, 0, :] = rad[:, 0, :] + np.asarray(rand_ini, dtype=np.float32)
    if mode == "ref32":
        tmp = np.mod(np.cumsum(rad, axis=1, dtype=np.float32), np.float32(1))
        over = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0
        shift = np.zeros_like(rad)
        shift[:, 1:, :] = over * np.float32(-1.0)
        phase = np.cumsum(rad + shift, axis=1, dtype=np.float32)
        sines = np.sin(phase * np.float32(2) * np.float32(np.pi)).astype(np.float64)
    else:
        r = rad.astype(np.float64)
        r = r - np.floor(r)
        # exact running sum mod 1: 2^-64 fixed point like the kernel (python ints are exact)
        fx = np.round(r * 2.0 ** 64).astype(object)   # exact: r has <= 24 significant bits
        phase = np.empty((B, S, H), dtype=np.float64)
        for b in range(B):
            for h in range(H):
                run = 0
                col = fx[b, :, h]
                out = phase[b, :, h]
                for i in range(S):
                    run = (run + int(col[i])) & ((1 << 64) - 1)
                    out[i] = (run >> 11) * 2.0 ** -53
        sines = np.sin(phase * 2.0 * np.pi)
    sine_waves = sines * sine_amp
    uv = (f0_up > 0).astype(np.float64)[:, :, None]
    noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3
    return sine_waves * uv + noise_amp * np.asarray(noise, dtype=np.float64)


def source_module(f0_up, sampling_rate, lin_w, lin_b, rand_ini, noise, mode="exact"):
    """SourceModuleHnNSF.forward (models.py:337-350): tanh(Linear(sine_waves)) -> [B,S]."""
    sw = sinegen(f0_up, sampling_rate, rand_ini, noise, mode=mode)
    return np.tanh(sw @ np.asarray(lin_w, dtype=np.float64).reshape(-1) + float(np.asarray(lin_b).reshape(-1)[0]))


def resblock1(sd, prefix, x, k, dilations):
    """ResBlock1.forward (models.py:103-110)."""
    for m, d in enumerate(dilations):
        xt = lrelu(x, LRELU_SLOPE)
        xt = conv1d(xt, sd[f"{prefix}convs1.{m}.weight"], sd[f"{prefix}convs1.{m}.bias"], dilation=d,
                    padding=get_padding(k, d))
        xt = lrelu(xt, LRELU_SLOPE)
        xt = conv1d(xt, sd[f"{prefix}convs2.{m}.weight"], sd[f"{prefix}convs2.{m}.bias"], dilation=1,
                    padding=get_padding(k, 1))
        x = xt + x
    return x


def resblock2_inplace(sd, prefix, x, k, dilations):
    """ResBlock2.forward (models.py:150-155) INCLUDING its aliasing: `xt = F.leaky_relu(x, LRELU_SLOPE, inplace=True)`
    rewrites x itself, so (a) the residual added is lrelu(x), not x, and (b) the caller's tensor -- the stage input that
    Generator.forward hands to all three ResBlocks (models.py:426-430) -- is left LeakyReLU'd once more after every block.
    Returns (block output, the stage input as the next block will see it)."""
    x = lrelu(x, LRELU_SLOPE)                 # in place on the shared stage input
    shared_after = x
    for m, d in enumerate(dilations):
        if m > 0:
            x = lrelu(x, LRELU_SLOPE)         # in place on the block's own intermediate
        xt = conv1d(x, sd[f"{prefix}convs.{m}.weight"], sd[f"{prefix}convs.{m}.bias"], dilation=d, padding=get_padding(k, d))
        x = xt + x
    return x, shared_after


def generator_forward(sd, h, mel, f0, rand_ini, noise, mode="exact", dtype=np.float64, return_source=False):
    """Generator.forward (models.py:407-438) with weight norm already folded (plain `weight` keys).
    mel [B,M,T], f0 [B,T]; returns wav [B,1,T*hop]."""
    sd = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    mel = np.asarray(mel, dtype=dtype)
    rates, ksz = h["upsample_rates"], h["upsample_kernel_sizes"]
    hop = int(np.prod(rates))
    f0_up = f0_upsample(f0, hop)
    har = source_module(f0_up, h["sampling_rate"], sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"],
                        rand_ini, noise, mode=mode).astype(dtype)[:, None, :]
    x = conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(rates, ksz)):
        x = lrelu(x, LRELU_SLOPE)
        x = conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], u, (k - u) // 2)
        if i + 1 < len(rates):
            stride_f0 = int(np.prod(rates[i + 1:]))
            xs_src = conv1d(har, sd[f"noise_convs.{i}.weight"], sd[f"noise_convs.{i}.bias"], stride=stride_f0,
                            padding=stride_f0 // 2)
        else:
            xs_src = conv1d(har, sd[f"noise_convs.{i}.weight"], sd[f"noise_convs.{i}.bias"])
        x = x + xs_src
        xs = None
        for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            if str(h.get("resblock", "1")) == "1":
                y = resblock1(sd, f"resblocks.{i * nk + j}.", x, rk, rd)
            else:
                y, x = resblock2_inplace(sd, f"resblocks.{i * nk + j}.", x, rk, rd)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = lrelu(x, 0.01)  # F.leaky_relu default slope (models.py:434)
    x = conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    x = np.tanh(x)
    if return_source:
        return x, har
    return x


def make_generator_weights(seed, h, scale=1.0):
    """Seeded synthetic generator weights, weight-norm already folded, reference key names/shapes
    (models.py:353-405).  std chosen so activations stay O(1) through the stack."""
    rng = np.random.RandomState(seed)
    sd = {}
    C0 = h["upsample_initial_channel"]
    M = h["num_mels"]

    def conv(name, co, ci, k, gain=1.0):
        sd[name + ".weight"] = (rng.randn(co, ci, k) * gain * scale / math.sqrt(ci * k)).astype(np.float32)
        sd[name + ".bias"] = (rng.randn(co) * 0.05).astype(np.float32)

    conv("conv_pre", C0, M, 7)
    rates, ksz = h["upsample_rates"], h["upsample_kernel_sizes"]
    for i, (u, k) in enumerate(zip(rates, ksz)):
        ci, co = C0 // (2 ** i), C0 // (2 ** (i + 1))
        sd[f"ups.{i}.weight"] = (rng.randn(ci, co, k) * scale * math.sqrt(u / (ci * k)) * 1.4).astype(np.float32)
        sd[f"ups.{i}.bias"] = (rng.randn(co) * 0.05).astype(np.float32)
        if i + 1 < len(rates):
            s = int(np.prod(rates[i + 1:]))
            conv(f"noise_convs.{i}", co, 1, 2 * s, gain=1.0)
        else:
            conv(f"noise_convs.{i}", co, 1, 1, gain=1.0)
        for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * len(h['resblock_kernel_sizes']) + j}."
            for m in range(len(rd)):
                conv(p + f"convs1.{m}", co, co, rk, gain=1.0)
                conv(p + f"convs2.{m}", co, co, rk, gain=0.5)
    conv("conv_post", 1, co, 7, gain=1.0)
    sd["m_source.l_linear.weight"] = (rng.randn(1, 9) * 0.5).astype(np.float32)
    sd["m_source.l_linear.bias"] = (rng.randn(1) * 0.1).astype(np.float32)
    return sd
