"""GPU test of the batched inference driver: bucketed / padded / masked batches equal the direct padded call."""
import numpy as np
import pytest
import torch

from fish_diffusion_b200 import BatchedSynthesizer, DIFFUSIONS, Generator, synthetic
from gpu_util import dev

pytestmark = pytest.mark.gpu


def test_batched_synthesizer_matches_direct_padded_batches(golden_cfg):
    wcfg = golden_cfg["WN_TC"]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **wcfg), mel_channels=64,
                                 sampler_interval=250, spec_min=[-5.0], spec_max=[0.0], noise_predictor="unipc")).to(dev())
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.wavenet_weights(1, **wcfg).items()})
    h = dict(golden_cfg["VOC_SMALL"], num_mels=64)
    gen = Generator(h).to(dev())
    gen.remove_weight_norm()
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.generator_weights(2, h).items()})
    g = torch.Generator().manual_seed(3)
    lengths = [300, 37, 256, 129, 90]
    feats = [torch.randn(L, 64, generator=g).to(dev()) for L in lengths]
    f0 = [(150 + 100 * torch.rand(L, generator=g)).to(dev()) for L in lengths]
    drv = BatchedSynthesizer(diff, gen, max_batch=2, bucket=128)
    torch.manual_seed(5)
    wavs, mels = drv(feats, f0, seed=11, return_mel=True)
    hop = 64
    for L, w, m in zip(lengths, wavs, mels):
        assert w.shape == (L * hop,) and m.shape == (L, 64)
        assert torch.isfinite(w).all() and torch.isfinite(m).all() and float(w.abs().max()) <= 1.0
    # batch 0 of the plan holds items [0, 2] padded to 384 frames: recompute it directly
    from fish_diffusion_b200 import plan_batches
    (idx, T), = plan_batches(lengths, 2, 128)[:1]
    assert idx == [0, 2] and T == 384
    feat = torch.zeros(2, T, 64, device=dev())
    mask = torch.ones(2, T, dtype=torch.bool, device=dev())
    for j, i in enumerate(idx):
        feat[j, :lengths[i]] = feats[i]
        mask[j, :lengths[i]] = False
    mel = diff(feat, x_masks=mask, cond_masks=mask, seed=11)
    assert torch.equal(mel[0, :300], mels[0]) and torch.equal(mel[1, :256], mels[2])
    # (padding frames carry sampler state with eps = 0 there, as in the reference; callers crop by length)


def test_sharded_batch_reproduces_unsharded_sampler_bitwise(golden_cfg):
    """batch sharding has no data-path collective and the Philox draws are indexed by the global element, so
    items [2,4) sampled alone (first_item=2) equal rows 2..3 of the 4-item run bit for bit (free-running noise)."""
    wcfg = golden_cfg["WN_TC"]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **wcfg), mel_channels=64,
                                 sampler_interval=100, spec_min=[-5.0], spec_max=[0.0], noise_predictor="naive")).to(dev())
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.wavenet_weights(1, **wcfg).items()})
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(4, 200, 64, generator=g).to(dev())
    full = diff(feats, seed=21)
    lo = diff(feats[:2].contiguous(), seed=21, first_item=0)
    hi = diff(feats[2:].contiguous(), seed=21, first_item=2)
    assert torch.equal(full[:2], lo) and torch.equal(full[2:], hi)
    assert not torch.equal(hi, diff(feats[2:].contiguous(), seed=21))       # without the offset the noise differs
