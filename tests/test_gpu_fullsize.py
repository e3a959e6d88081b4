"""GPU parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run B=32 x T=4000):
batch-item independence (what makes the batch-shard multi-GPU path exact), tensor-core == SIMT twin on a slice,
and sampler invariance to how the batch is split."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import DIFFUSIONS, Generator, WaveNet, synthetic
from gpu_util import dev

pytestmark = pytest.mark.gpu

WN_FULL = dict(mel_channels=128, d_encoder=256, residual_channels=512, residual_layers=20, use_linear_bias=True,
               dilation_cycle=4)


@pytest.fixture(scope="module")
def full_net():
    net = WaveNet(**WN_FULL).to(dev())
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.wavenet_weights(0, **WN_FULL).items()})
    return net.eval()


def test_denoiser_full_batch_item_independence(full_net):
    """One evaluation at B=32, T=4000 (config #2 shape): every item equals the same item evaluated alone."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(32, 128, 4000, generator=g).to(dev())
    c = torch.randn(32, 256, 4000, generator=g).to(dev())
    t = torch.tensor([990], device=dev())
    with torch.no_grad():
        y = full_net(x, t, c)
        assert torch.isfinite(y).all()
        for i in (0, 13, 31):
            yi = full_net(x[i:i + 1].contiguous(), t, c[i:i + 1].contiguous())
            assert torch.equal(yi[0], y[i]), i          # bit-identical: tiles never mix items


def test_denoiser_full_width_tc_equals_simt_on_slice(full_net):
    g = torch.Generator().manual_seed(8)
    x = torch.randn(1, 128, 4000, generator=g).to(dev())
    c = torch.randn(1, 256, 4000, generator=g).to(dev())
    t = torch.tensor([500], device=dev())
    simt = WaveNet(**WN_FULL, backend="simt").to(dev())
    simt.load_state_dict(full_net.state_dict())
    with torch.no_grad():
        a, b = full_net(x, t, c), simt.eval()(x, t, c)
    e = rel_l2(a.cpu().numpy(), b.cpu().numpy())
    print(f"full-width denoiser tc vs simt rel-L2 {e:.2e}")
    assert e < 2e-5


def test_time_tiling_does_not_leak_across_ragged_T(full_net):
    """T not a multiple of the 128-row tile: the valid part equals a zero-padded-equivalent run (conv zero padding
    at the right edge is what the TMA out-of-bounds fill provides)."""
    g = torch.Generator().manual_seed(9)
    T = 1000
    x = torch.randn(2, 128, T, generator=g).to(dev())
    c = torch.randn(2, 256, T, generator=g).to(dev())
    t = torch.tensor([10, 700], device=dev())
    with torch.no_grad():
        y = full_net(x, t, c)
        # receptive field is 75 frames per side: frames < T-75-... of a longer, different continuation must agree
        x2 = torch.cat([x, torch.randn(2, 128, 300, generator=g).to(dev())], dim=2)
        c2 = torch.cat([c, torch.randn(2, 256, 300, generator=g).to(dev())], dim=2)
        y2 = full_net(x2, t, c2)
    assert torch.allclose(y[:, :, :T - 76], y2[:, :, :T - 76], rtol=0, atol=2e-5)
    assert not torch.allclose(y[:, :, T - 5:], y2[:, :, T - 5:T], atol=1e-3)


def test_sampler_invariant_to_batch_split():
    """The naive sampler with injected noise: running items [0:2] and [2:4] separately equals the 4-item batch."""
    cfg = dict(WN_FULL, residual_layers=4)
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **cfg), mel_channels=128,
                                 sampler_interval=250, spec_min=[-5.0], spec_max=[0.0], noise_predictor="naive")).to(dev())
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.wavenet_weights(3, **cfg).items()})
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(4, 600, 256, generator=g).to(dev())
    xT = torch.randn(4, 128, 600, generator=g).to(dev())
    nz = [torch.randn(4, 128, 600, generator=g).to(dev()) for _ in range(4)]
    full = diff(feats, x_T=xT, step_noises=nz)
    lo = diff(feats[:2].contiguous(), x_T=xT[:2].contiguous(), step_noises=[n[:2].contiguous() for n in nz])
    hi = diff(feats[2:].contiguous(), x_T=xT[2:].contiguous(), step_noises=[n[2:].contiguous() for n in nz])
    assert torch.equal(full, torch.cat([lo, hi]))


def test_vocoder_long_item_independence_and_bounds():
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "nsf_configs", "config_v1.json")) as f:
        h = json.load(f)
    gen = Generator(h).to(dev())
    gen.remove_weight_norm()
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.generator_weights(5, h).items()})
    g = torch.Generator().manual_seed(12)
    B, T = 3, 500                                    # 256 000 samples per item
    mel = (torch.randn(B, 128, T, generator=g) - 2.5).clamp(-11.5, 2).to(dev())
    f0 = (200.0 + 50 * torch.sin(torch.arange(T) / 30.0)).repeat(B, 1)
    f0[:, 100:140] = 0
    f0 = f0.to(dev())
    ri = torch.rand(B, 9, generator=g).to(dev())
    nz = torch.randn(B, T * 512, 9, generator=g).to(dev())
    wav = gen(mel, f0, rand_ini=ri, sine_noise=nz)
    assert wav.shape == (B, 1, T * 512) and torch.isfinite(wav).all() and float(wav.abs().max()) <= 1.0
    one = gen(mel[1:2].contiguous(), f0[1:2].contiguous(), rand_ini=ri[1:2].contiguous(), sine_noise=nz[1:2].contiguous())
    assert torch.equal(one[0], wav[1])
