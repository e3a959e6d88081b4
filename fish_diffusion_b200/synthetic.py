"""Seeded synthetic weights with the reference's state_dict keys/shapes (no checkpoint is reachable offline).
Used by bench.py and examples; numpy RNG so any box regenerates the same tensors.  Magnitudes follow the reference
initialisers (kaiming-normal convs wavenet.py:75, xavier-uniform linears wavenet.py:37); WaveNet.output_projection is
random rather than zero-initialised so that outputs depend on the inputs."""
import math

import numpy as np


def wavenet_weights(seed, mel_channels=128, d_encoder=256, residual_channels=512, residual_layers=20,
                    use_linear_bias=True, **_unused):
    rng = np.random.RandomState(seed)
    M, E, C = mel_channels, d_encoder, residual_channels
    sd = {}

    def conv(name, co, ci, k):
        sd[name + ".conv.weight"] = (rng.randn(co, ci, k) * math.sqrt(2.0 / (ci * k))).astype(np.float32)
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


def generator_weights(seed, h):
    """NSF-HiFiGAN generator weights, weight norm already folded (plain `weight` keys)."""
    rng = np.random.RandomState(seed)
    sd = {}
    C0, M = h["upsample_initial_channel"], h["num_mels"]

    def conv(name, co, ci, k, gain=1.0):
        sd[name + ".weight"] = (rng.randn(co, ci, k) * gain / math.sqrt(ci * k)).astype(np.float32)
        sd[name + ".bias"] = (rng.randn(co) * 0.05).astype(np.float32)

    conv("conv_pre", C0, M, 7)
    rates, ksz = h["upsample_rates"], h["upsample_kernel_sizes"]
    nk = len(h["resblock_kernel_sizes"])
    co = C0
    for i, (u, k) in enumerate(zip(rates, ksz)):
        ci, co = C0 // (2 ** i), C0 // (2 ** (i + 1))
        sd[f"ups.{i}.weight"] = (rng.randn(ci, co, k) * math.sqrt(u / (ci * k)) * 1.4).astype(np.float32)
        sd[f"ups.{i}.bias"] = (rng.randn(co) * 0.05).astype(np.float32)
        if i + 1 < len(rates):
            conv(f"noise_convs.{i}", co, 1, 2 * int(np.prod(rates[i + 1:])))
        else:
            conv(f"noise_convs.{i}", co, 1, 1)
        for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}."
            if str(h.get("resblock", "1")) == "1":
                for m in range(len(rd)):
                    conv(p + f"convs1.{m}", co, co, rk)
                    conv(p + f"convs2.{m}", co, co, rk, gain=0.5)
            else:
                for m in range(len(rd)):
                    conv(p + f"convs.{m}", co, co, rk, gain=0.7)
    conv("conv_post", 1, co, 7)
    sd["m_source.l_linear.weight"] = (rng.randn(1, 9) * 0.5).astype(np.float32)
    sd["m_source.l_linear.bias"] = (rng.randn(1) * 0.1).astype(np.float32)
    return sd
