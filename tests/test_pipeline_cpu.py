"""CPU tests of the 'next' rows N1-N3: batching plan, DiffSinger feature fusion, Lightning-checkpoint key handling."""
import numpy as np
import pytest
import torch

from fish_diffusion_b200 import DiffSinger, load_checkpoint, pitch_to_scale, plan_batches
from fish_diffusion_b200.pipeline import padding_waste

MODEL_CFG = dict(
    text_encoder=dict(type="NaiveProjectionEncoder", input_size=24, output_size=32),
    speaker_encoder=dict(type="NaiveProjectionEncoder", input_size=10, output_size=32, use_embedding=True),
    pitch_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=32, preprocessing=pitch_to_scale),
    diffusion=dict(type="GaussianDiffusion", mel_channels=16, sampler_interval=10, spec_min=[-5.0], spec_max=[0.0],
                   denoiser=dict(type="WaveNetDenoiser", mel_channels=16, d_encoder=32, residual_channels=64,
                                 residual_layers=2, use_linear_bias=True)),
)


def test_plan_batches_covers_everything_once():
    rng = np.random.RandomState(0)
    lengths = rng.randint(1, 4000, size=137).tolist()
    batches = plan_batches(lengths, max_batch=32, bucket=128, max_frames=32 * 4096)
    seen = sorted(i for idx, _ in batches for i in idx)
    assert seen == list(range(137))
    for idx, T in batches:
        assert 1 <= len(idx) <= 32 and T % 128 == 0
        assert all(lengths[i] <= T for i in idx) and max(lengths[i] for i in idx) > T - 128
        assert len(idx) * T <= 32 * 4096 or len(idx) == 1
    assert padding_waste(lengths, batches) < 0.12          # length-sorted buckets keep the padding small
    with pytest.raises(ValueError):
        plan_batches([5, 0])


def test_diffsinger_forward_features_matches_definition():
    torch.manual_seed(0)
    m = DiffSinger(MODEL_CFG)
    B, T = 3, 11
    contents = torch.randn(B, T, 24)
    speakers = torch.tensor([1, 4, 9])
    pitches = torch.rand(B, T) * 900 + 60
    lens = torch.tensor([11, 7, 3])
    out = m.forward_features(speakers=speakers, contents=contents, contents_lens=lens, contents_max_len=T, mel_lens=lens,
                             mel_max_len=T, pitches=pitches)
    te, se, pe = m.text_encoder.projection, m.speaker_encoder.embedding, m.pitch_encoder.projection
    scale = ((pitches - 50.0) / 1050.0).clamp(0, 1)[..., None]
    want = contents @ te.weight.T + te.bias + se.weight[speakers][:, None] + scale @ pe.weight.T + pe.bias
    assert torch.allclose(out["features"], want, atol=1e-6)
    assert out["features"].shape == (B, T, 32)
    assert torch.equal(out["x_masks"], torch.arange(T)[None] >= lens[:, None])
    assert out["cond_masks"] is out["x_masks"]
    # reference parameter names (model.* keys of a Lightning checkpoint)
    keys = set(m.state_dict())
    assert {"text_encoder.projection.weight", "speaker_encoder.embedding.weight", "pitch_encoder.projection.bias",
            "diffusion.denoise_fn.input_projection.conv.weight", "diffusion.naive_noise_predictor.clip_min",
            "diffusion.spec_min"} <= keys


def test_load_checkpoint_prefix_and_vocoder_keys(tmp_path):
    torch.manual_seed(1)
    src, dst = DiffSinger(MODEL_CFG), DiffSinger(MODEL_CFG)
    sd = {"model." + k: v for k, v in src.state_dict().items()}
    sd["vocoder.model.conv_pre.weight"] = torch.zeros(3)          # dropped like utils/inference.py:18-22
    torch.save({"state_dict": sd}, tmp_path / "a.ckpt")
    missing, unexpected = load_checkpoint(dst, str(tmp_path / "a.ckpt"), device="cpu")
    assert not missing and not unexpected
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k


def test_time_folded_conv_weight_equals_the_conv():
    """Host logic of the vocoder's time folding (nsf_hifigan.fold_conv_weight): the block-Toeplitz tap-GEMM on the
    [T/F, F*C] view reproduces torch's dilated 'same' Conv1d exactly (float64), for every (K, d, F) the packer uses."""
    import torch
    from fish_diffusion_b200.nsf_hifigan import Generator, fold_conv_weight
    g = torch.Generator().manual_seed(0)
    for C, K, d in [(16, 3, 1), (16, 7, 3), (16, 11, 5), (32, 11, 1), (32, 7, 1), (64, 11, 1), (16, 11, 1)]:
        F = Generator._fold_factor(C, C, K, d)
        assert F > 1 and C * F == 128
        w = torch.randn(C, C, K, generator=g, dtype=torch.float64)
        x = torch.randn(2, C, 40 * F, generator=g, dtype=torch.float64)          # [B, C, T]
        ref = torch.nn.functional.conv1d(x, w, padding=(K * d - d) // 2, dilation=d)
        wf, srows = fold_conv_weight(w, d, F)
        assert len(srows) <= 16
        xf = x.transpose(1, 2).reshape(2, 40, F * C)                             # [B, T/F, F*C]: same memory as [B,T,C]
        yf = torch.zeros(2, 40, F * C, dtype=torch.float64)
        for si, s in enumerate(srows):
            seg = torch.zeros_like(xf)
            lo, hi = max(0, -s), min(40, 40 - s)
            if hi > lo:
                seg[:, lo:hi] = xf[:, lo + s:hi + s]
            yf += seg @ wf[:, si * F * C:(si + 1) * F * C].T
        got = yf.reshape(2, 40 * F, C).transpose(1, 2)
        assert torch.allclose(got, ref, atol=1e-12), (C, K, d, float((got - ref).abs().max()))
    # convs the packer leaves alone
    assert Generator._fold_factor(128, 128, 11, 1) == 1 and Generator._fold_factor(32, 32, 11, 3) == 1
    assert Generator._fold_factor(64, 64, 3, 1) == 1 and Generator._fold_factor(16, 32, 3, 1) == 1


def test_folded_pair_pack_equals_the_conv():
    """Generator._pack_conv_folded_pair: the C=16 stage runs the fused pair kernel on the time-folded view [T/2, 32] as a
    DILATION-1 conv with K' = 2*max|shift|+1 taps of 32x32 blocks.  Exactness of that rewrite, on CPU with torch convs."""
    import torch
    import torch.nn.functional as F
    from fish_diffusion_b200.nsf_hifigan import fold_conv_weight
    torch.manual_seed(0)
    C, T, Fd = 16, 64, 2
    for k, d in ((3, 1), (3, 5), (7, 3), (11, 1), (11, 5)):
        w = torch.randn(C, C, k, dtype=torch.float64)
        x = torch.randn(1, C, T, dtype=torch.float64)
        y = F.conv1d(x, w, padding=(k - 1) // 2 * d, dilation=d)                      # [1,C,T]
        wf, srows = fold_conv_weight(w, d, Fd)                                        # [F*C, S*F*C]
        hmax = max(abs(srows[0]), abs(srows[-1]))
        Kp = 2 * hmax + 1
        W = torch.zeros((Fd * C, Kp, Fd * C), dtype=torch.float64)
        wf3 = wf.reshape(Fd * C, len(srows), Fd * C)
        for j, sr in enumerate(srows):
            W[:, sr + hmax, :] = wf3[:, j, :]
        # folded view: row r holds time steps 2r, 2r+1 -> channels (f, c)
        xf = x[0].t().reshape(T // Fd, Fd * C).t()[None]                              # [1, F*C, T/F]
        yf = F.conv1d(xf, W.permute(0, 2, 1).contiguous(), padding=hmax)              # [1, F*C, T/F]
        back = yf[0].t().reshape(T, C).t()[None]
        assert torch.allclose(back, y, atol=1e-10), (k, d)
        assert Kp % 2 == 1 and (Kp - 1) <= 56
