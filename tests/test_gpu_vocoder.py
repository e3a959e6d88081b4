"""GPU parity: NSF-HiFiGAN source module / generator and the mel front end against the reference's golden vectors
and the float64 oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import Generator, PitchAdjustableMelSpectrogram
from gpu_util import dev
from oracle import mel as omel
from oracle import nsf_hifigan as ovoc

pytestmark = pytest.mark.gpu


def T_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _ri(raw):
    r = raw.copy()
    r[:, 0] = 0
    return r


def _gen(h, sd, **kw):
    gen = Generator(h, **kw).to(dev())
    gen.remove_weight_norm()
    r = gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    return gen.eval()


def test_source_module_vs_reference_and_oracle(golden, golden_cfg):
    g = golden("vocoder")
    h = golden_cfg["VOC_SMALL"]
    sd = ovoc.make_generator_weights(31, h)
    gen = _gen(h, sd)
    har = gen.source(T_(g["src_f0"]), 64, rand_ini=T_(g["src_rand_ini_raw"]), sine_noise=T_(g["src_sine_noise"]))
    har = har.cpu().numpy()
    oracle = ovoc.source_module(ovoc.f0_upsample(g["src_f0"], 64), 44100, sd["m_source.l_linear.weight"],
                                sd["m_source.l_linear.bias"], _ri(g["src_rand_ini_raw"]), g["src_sine_noise"],
                                mode="exact")
    e_or, e_ref = np.abs(har - oracle).max(), np.abs(har - g["src_har"]).max()
    print(f"source module: max|cuda-oracle| {e_or:.2e}, max|cuda-reference| {e_ref:.2e}")
    assert e_or < 2e-6           # exact-phase kernel == exact-phase oracle
    assert e_ref < 2e-5          # reference's own fp32 phase noise
    assert np.all(np.abs(har) <= 1.0)


def test_source_module_long_sequence_phase_is_exact():
    """Size-independent property: the phase scan is exact integer arithmetic, so the excitation of a constant-f0
    item is periodic with the period predicted from rad = fp32(f0/sr) no matter how long the signal is."""
    h = dict(resblock="1", upsample_rates=[8, 8, 2, 2, 2], upsample_kernel_sizes=[16, 16, 8, 2, 2],
             upsample_initial_channel=32, resblock_kernel_sizes=[3], resblock_dilation_sizes=[[1, 3, 5]], num_mels=16,
             hop_size=512, sampling_rate=44100)
    gen = Generator(h).to(dev())
    T = 2000                                         # 1 024 000 samples
    f0 = torch.full((1, T), 441.0, device=dev())
    ri = torch.zeros(1, 9, device=dev())
    nz = torch.zeros(1, T * 512, 9, device=dev())
    har = gen.source(f0, 512, rand_ini=ri, sine_noise=nz).cpu().numpy()[0]
    lw = gen.m_source.l_linear.weight.detach().cpu().numpy().reshape(-1).astype(np.float64)
    lb = float(gen.m_source.l_linear.bias.detach().cpu())
    f0_up = ovoc.f0_upsample(np.full((1, T), 441.0, dtype=np.float32), 512)[0]      # bit-exact vs ATen
    want = lb
    for hh in range(1, 10):
        rad = np.mod((f0_up * np.float32(hh)) / np.float32(44100.0), np.float32(1)).astype(np.float64)
        phase = np.cumsum(rad) % 1.0           # float64 running sum: exact to ~1e-12 over 1M samples
        want = want + lw[hh - 1] * 0.1 * np.sin(2 * np.pi * phase)
    want = np.tanh(want)
    assert np.abs(har - want).max() < 5e-6


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_generator_vs_reference(golden, golden_cfg, backend):
    g = golden("vocoder")
    h = golden_cfg["VOC_SMALL"]
    sd = ovoc.make_generator_weights(31, h)
    gen = _gen(h, sd, backend=backend)
    wav = gen(T_(g["voc_small_mel"]), T_(g["voc_small_f0"]), rand_ini=T_(g["voc_small_rand_ini_raw"]),
              sine_noise=T_(g["voc_small_sine_noise"])).cpu().numpy()
    e = rel_l2(wav, g["voc_small_wav"])
    print(f"generator[{backend}] rel-L2 vs reference {e:.2e}")
    assert wav.shape == g["voc_small_wav"].shape
    assert e < 1e-4


@pytest.mark.parametrize("name", ["config_v1", "config_v1_256"])
def test_generator_real_configs_vs_oracle(name):
    """The two shipped JSON configs (hop 512 / hop 256) at full channel widths, short T, against the fp64 oracle."""
    here = os.path.join(os.path.dirname(__file__), "golden", "nsf_configs")
    with open(os.path.join(here, name + ".json")) as f:
        h = json.load(f)
    sd = ovoc.make_generator_weights(77, h)
    rng = np.random.RandomState(78)
    B, T = 1, 6
    hop = int(np.prod(h["upsample_rates"]))
    mel = (rng.randn(B, 128, T) - 2.5).clip(-11.5, 2).astype(np.float32)
    f0 = np.array([[220.0, 230.0, 0.0, 0.0, 300.0, 310.0]], dtype=np.float32)
    ri = rng.rand(B, 9).astype(np.float32)
    nz = rng.randn(B, T * hop, 9).astype(np.float32)
    ref = ovoc.generator_forward(sd, h, mel, f0, _ri(ri), nz, mode="exact")
    for backend in ("auto", "simt"):
        gen = _gen(h, sd, backend=backend)
        wav = gen(T_(mel), T_(f0), rand_ini=T_(ri), sine_noise=T_(nz)).cpu().numpy()
        e = rel_l2(wav, ref)
        print(f"generator[{name},{backend}] rel-L2 vs oracle {e:.2e}")
        assert e < 5e-5


@pytest.mark.parametrize("ks", [0, 5, -5])
def test_mel_front_end_vs_reference(golden, ks):
    g = golden("mel")
    for backend in ("auto", "simt"):
        pam = PitchAdjustableMelSpectrogram(backend=backend)
        spec = pam(T_(g["mel_wav"]), key_shift=ks).cpu().numpy()
        ref = g[f"mel_spec_ks{ks}"]
        e = rel_l2(spec, ref)
        print(f"mel[ks={ks},{backend}] rel-L2 vs reference {e:.2e}")
        assert spec.shape == ref.shape
        assert e < 5e-5


def test_mel_speed_and_log(golden):
    from fish_diffusion_b200 import dynamic_range_compression
    g = golden("mel")
    pam = PitchAdjustableMelSpectrogram()
    spec = pam(T_(g["mel_wav"]), speed=0.5)
    assert rel_l2(spec.cpu().numpy(), g["mel_spec_speed"]) < 5e-5
    lg = dynamic_range_compression(spec).cpu().numpy()
    assert rel_l2(lg, omel.dynamic_range_compression(g["mel_spec_speed"].astype(np.float64))) < 1e-5
