"""Oracle: WaveNet denoiser (reference fish_diffusion/modules/wavenet.py).  TEST INFRASTRUCTURE ONLY.

Weights are passed as a dict of numpy arrays keyed exactly like the reference ``state_dict``
(SURVEY.md section 8b), e.g. ``residual_layers.3.conv_layer.conv.weight``.
"""
import math

import numpy as np


def conv1d(x, w, b=None, dilation=1, padding=0, stride=1):
    """torch.nn.functional.conv1d restated: x [B,Ci,T], w [Co,Ci,K] -> [B,Co,To] (zero padding)."""
    B, Ci, T = x.shape
    Co, Ci2, K = w.shape
    assert Ci == Ci2
    xp = np.zeros((B, Ci, T + 2 * padding), dtype=x.dtype)
    xp[:, :, padding:padding + T] = x
    To = (T + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    out = np.empty((B, Co, To), dtype=x.dtype)
    # one BLAS GEMM per item: W[Co, K*Ci] @ im2col[K*Ci, To]  (tap-major columns)
    w2 = np.ascontiguousarray(np.transpose(w, (0, 2, 1)).reshape(Co, K * Ci))
    for bi in range(B):
        cols = np.concatenate([xp[bi, :, k * dilation:k * dilation + (To - 1) * stride + 1:stride] for k in range(K)],
                              axis=0)
        out[bi] = w2 @ cols
    if b is not None:
        out += b[None, :, None]
    return out


def mish(x):
    """wavenet.py:8-10  x * tanh(softplus(x)); softplus uses torch's threshold=20 linearisation."""
    sp = np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))
    return x * np.tanh(sp)


def diffusion_embedding(steps, dim, dtype=np.float64):
    """wavenet.py:20-27.  The frequency table is float32 whatever the input dtype: `torch.arange(half_dim) * -emb`
    is int64 x python-float -> float32 (torch default dtype), and so is its exp; only the product with the step
    (and sin/cos) runs in the step's dtype."""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    arg = np.arange(half).astype(np.float32) * np.float32(-emb)
    table = np.exp(arg.astype(np.float64)).astype(np.float32)
    emb = np.asarray(steps, dtype=dtype)[:, None] * table[None, :].astype(dtype)
    return np.concatenate([np.sin(emb), np.cos(emb)], axis=-1)


def linear(x, w, b=None):
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def residual_block(sd, prefix, x, conditioner, step, dilation):
    """wavenet.py:106-120 (ResidualBlock.forward)."""
    g = lambda k: sd.get(prefix + k)
    d = linear(step, g("diffusion_projection.linear.weight"), g("diffusion_projection.linear.bias"))[:, :, None]
    c = conv1d(conditioner, g("conditioner_projection.conv.weight"), g("conditioner_projection.conv.bias"))
    y = x + d
    y = conv1d(y, g("conv_layer.conv.weight"), g("conv_layer.conv.bias"), dilation=dilation, padding=dilation) + c
    C = x.shape[1]
    gate, filt = y[:, :C], y[:, C:]
    y = 1.0 / (1.0 + np.exp(-gate)) * np.tanh(filt)
    y = conv1d(y, g("output_projection.conv.weight"), g("output_projection.conv.bias"))
    residual, skip = y[:, :C], y[:, C:]
    return (x + residual) / math.sqrt(2.0), skip


def wavenet_forward(sd, x, diffusion_step, conditioner, x_masks=None, cond_masks=None, dilation_cycle=None,
                    dtype=np.float64):
    """wavenet.py:194-236 (WaveNet.forward).  x [B,M,T] (or [B,1,M,T]), diffusion_step [B] or [1],
    conditioner [B,E,T], masks [B,T] bool (True = masked)."""
    sd = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    x = np.asarray(x, dtype=dtype)
    conditioner = np.asarray(conditioner, dtype=dtype)
    use_4 = x.ndim == 4
    if use_4:
        x = x[:, 0]
    assert x.ndim == 3
    n_layers = len({k.split(".")[1] for k in sd if k.startswith("residual_layers.")})
    C = sd["input_projection.conv.weight"].shape[0]
    x = conv1d(x, sd["input_projection.conv.weight"], sd["input_projection.conv.bias"])
    x = np.maximum(x, 0.0)
    step = diffusion_embedding(np.asarray(diffusion_step, dtype=dtype), C, dtype)
    step = linear(step, sd["mlp.0.linear.weight"], sd.get("mlp.0.linear.bias"))
    step = mish(step)
    step = linear(step, sd["mlp.2.linear.weight"], sd.get("mlp.2.linear.bias"))
    if x_masks is not None:
        x = np.where(x_masks[:, None], 0.0, x)
    if cond_masks is not None:
        conditioner = np.where(cond_masks[:, None], 0.0, conditioner)
    skips = []
    for i in range(n_layers):
        dil = 2 ** (i % dilation_cycle) if dilation_cycle else 1
        x, s = residual_block(sd, f"residual_layers.{i}.", x, conditioner, step, dil)
        skips.append(s)
    x = np.sum(np.stack(skips), axis=0) / math.sqrt(n_layers)
    x = conv1d(x, sd["skip_projection.conv.weight"], sd["skip_projection.conv.bias"])
    x = np.maximum(x, 0.0)
    x = conv1d(x, sd["output_projection.conv.weight"], sd["output_projection.conv.bias"])
    if x_masks is not None:
        x = np.where(x_masks[:, None], 0.0, x)
    return x[:, None] if use_4 else x


def make_wavenet_weights(seed, mel_channels=128, d_encoder=256, residual_channels=512, residual_layers=20,
                         use_linear_bias=True, scale=1.0):
    """Seeded synthetic weights with the reference's key names and shapes (numpy RNG so that the same weights
    can be regenerated on any box).  output_projection is random, NOT zero.  Magnitudes follow
    the reference initialisers (kaiming-normal convs wavenet.py:75, xavier-uniform linears wavenet.py:37)."""
    rng = np.random.RandomState(seed)
    M, E, C = mel_channels, d_encoder, residual_channels
    sd = {}

    def conv(name, co, ci, k):
        sd[name + ".conv.weight"] = (rng.randn(co, ci, k) * math.sqrt(2.0 / (ci * k)) * scale).astype(np.float32)
        sd[name + ".conv.bias"] = (rng.uniform(-1, 1, co) / math.sqrt(ci * k)).astype(np.float32)

    def lin(name, co, ci, bias):
        a = math.sqrt(6.0 / (ci + co))
        sd[name + ".linear.weight"] = rng.uniform(-a, a, (co, ci)).astype(np.float32)
        if bias:
            sd[name + ".linear.bias"] = (rng.uniform(-1, 1, co) * 0.1).astype(np.float32)

    conv("input_projection", C, M, 1)
    lin("mlp.0", 4 * C, C, use_linear_bias)
    lin("mlp.2", C, 4 * C, use_linear_bias)
    for i in range(residual_layers):
        p = f"residual_layers.{i}."
        conv(p + "conv_layer", 2 * C, C, 3)
        lin(p + "diffusion_projection", C, C, use_linear_bias)
        conv(p + "conditioner_projection", 2 * C, E, 1)
        conv(p + "output_projection", 2 * C, C, 1)
    conv("skip_projection", C, C, 1)
    conv("output_projection", M, C, 1)
    return sd
