"""GPU parity against the round-2 golden vectors (tests/golden/make_golden_r2.py, generated from the UNMODIFIED reference):
100-step sampler trajectories through the tcgen05 back end, full-width training gradients, masks + input gradient
under grad, both shipped generator configs at T=128, the torchaudio-variant mel front end, checkpoints written by the
reference classes."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from fish_diffusion_b200 import DIFFUSIONS, Generator, NsfHifiGAN, WaveNet, formats, get_mel_from_audio, get_mel_transform
from fish_diffusion_b200 import dynamic_range_compression
from gpu_util import dev
from oracle import nsf_hifigan as ovoc
from oracle import wavenet as ownet

pytestmark = pytest.mark.gpu


def T_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def wn_weights(seed, cfg):
    return ownet.make_wavenet_weights(seed, **{k: v for k, v in cfg.items() if k != "dilation_cycle"})


def redraw(seed, draws, kinds):
    """The reference's random draws again: np.random.RandomState(seed), same order, same shapes (make_golden.RecordedRandom)."""
    rng = np.random.RandomState(int(seed))
    out = []
    for shape, kind in zip(draws, kinds):
        shape = [int(s) for s in shape if s > 0]
        out.append((rng.rand(*shape) if str(kind) == "rand" else rng.randn(*shape)).astype(np.float32))
    return out


# ------------------------------------------------------------------ 100-step trajectories, tcgen05 back end
@pytest.mark.parametrize("pred", ["naive", "unipc"])
@pytest.mark.parametrize("precision,tol", [("f16", 2e-4), ("f16x1", 5e-2)])
def test_100_step_trajectory_tc(golden, golden_cfg, pred, precision, tol):
    g = golden("r2_traj")
    cfg = golden_cfg["WN_TC"]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", backend="tc", precision=precision, **cfg),
                                 mel_channels=cfg["mel_channels"], noise_schedule="linear", timesteps=1000, max_beta=0.01,
                                 noise_loss="smoothed-l1", sampler_interval=10, spec_min=[-5.0], spec_max=[0.0],
                                 noise_predictor=pred)).to(dev()).eval()
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in wn_weights(61, cfg).items()})
    noises = [T_(a) for a in redraw(g[f"traj_{pred}_seed"], g[f"traj_{pred}_draws"], g[f"traj_{pred}_kinds"])]
    assert len(noises) == (101 if pred == "naive" else 1)
    mel = diff(T_(g["traj_features"]), sampler_interval=10, noise_predictor=pred, x_T=noises[0], step_noises=noises[1:])
    ref = g[f"traj_{pred}_mel"]
    e = rel_l2(mel.cpu().numpy(), ref)
    print(f"100-step {pred} [{precision}, tc] rel-L2 vs the reference trajectory: {e:.2e}")
    assert mel.shape == ref.shape and e < tol


# ------------------------------------------------------------------ full-width training gradients
def test_full_width_training_gradients(golden, golden_cfg):
    g = golden("r2_train_full")
    cfg = golden_cfg["WN_FULL"]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", backend="tc", **cfg),
                                 mel_channels=cfg["mel_channels"], noise_loss="smoothed-l1", sampler_interval=10,
                                 spec_min=[-5.0], spec_max=[0.0])).to(dev()).train()
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in wn_weights(71, cfg).items()})
    feats = T_(g["tf_features"]).requires_grad_(True)
    out = diff.train_step(feats, T_(g["tf_mel"]), t=T_(g["tf_t"]), noise=T_(g["tf_noise"]))
    assert abs(float(out["loss"]) - float(g["tf_loss"])) < 2e-5 * abs(float(g["tf_loss"]))
    assert rel_l2(out["epsilon"].detach().cpu().numpy(), g["tf_eps"]) < 2e-5
    out["loss"].backward()
    assert rel_l2(feats.grad.cpu().numpy(), g["tf_gfeatures"]) < 2e-4
    worst = dict(ref=("", 0.0), f64=("", 0.0), ref_vs_f64=("", 0.0), norm=("", 0.0))
    for k, p in diff.denoise_fn.named_parameters():
        gr = p.grad.detach().reshape(-1).cpu().numpy()
        idx = g[f"tf_g_{k}_idx"]
        got = gr[idx]
        e_ref = rel_l2(got, g[f"tf_g_{k}_val"])                 # vs the reference's fp32 autograd
        e_64 = rel_l2(got, g[f"tf64_g_{k}_val"])                # vs the same module in float64
        e_rr = rel_l2(g[f"tf_g_{k}_val"], g[f"tf64_g_{k}_val"])  # the reference's own fp32 noise
        e_n = abs(float(np.linalg.norm(gr.astype(np.float64))) - float(g[f"tf_g_{k}_norm"])) / float(g[f"tf_g_{k}_norm"])
        for name, e in (("ref", e_ref), ("f64", e_64), ("ref_vs_f64", e_rr), ("norm", e_n)):
            if e > worst[name][1]:
                worst[name] = (k, e)
        assert e_ref < 5e-4 and e_n < 5e-4, (k, e_ref, e_n)
    print("full-width gradients (C=512, L=20, B=1, T=128), worst parameter: "
          + ", ".join(f"{n} {v[1]:.2e} ({v[0]})" for n, v in worst.items()))


def test_masked_forward_and_gradients(golden, golden_cfg):
    g = golden("r2_train_masked")
    cfg = golden_cfg["WN_TC"]
    for backend in ("tc", "simt"):
        net = WaveNet(**cfg, backend=backend).to(dev())
        net.load_state_dict({k: torch.from_numpy(v) for k, v in wn_weights(81, cfg).items()})
        x = T_(g["tm_x"]).requires_grad_(True)
        c = T_(g["tm_c"]).requires_grad_(True)
        m = T_(g["tm_masks"])
        y = net(x, T_(g["tm_steps"]), c, x_masks=m, cond_masks=m)
        assert rel_l2(y.detach().cpu().numpy(), g["tm_y"]) < 2e-5
        ((y * T_(g["tm_w"])).sum() / y.numel()).backward()
        assert rel_l2(x.grad.cpu().numpy(), g["tm_gx"]) < 2e-4
        assert rel_l2(c.grad.cpu().numpy(), g["tm_gc"]) < 2e-4
        worst = 0.0
        for k, p in net.named_parameters():
            e = rel_l2(p.grad.cpu().numpy(), g[f"tm_g_{k}"])
            worst = max(worst, e)
            assert e < 2e-4, (k, e)
        print(f"masked forward/backward [{backend}]: worst parameter gradient rel-L2 {worst:.2e}")


# ------------------------------------------------------------------ generator, shipped configs, T = 128
@pytest.mark.parametrize("name", ["config_v1", "config_v1_256"])
def test_generator_shipped_configs_t128_vs_reference(golden, name):
    g = golden("r2_voc")
    with open(os.path.join(GOLDEN, "nsf_configs", name + ".json")) as f:
        h = json.load(f)
    sd = ovoc.make_generator_weights(int(g[f"voc_{name}_wseed"]), h)
    mel, f0 = g[f"voc_{name}_mel"], g[f"voc_{name}_f0"]
    B, T = f0.shape
    S = T * int(np.prod(h["upsample_rates"]))
    rng = np.random.RandomState(int(g[f"voc_{name}_rseed"]))
    ri = rng.rand(B, 9).astype(np.float32)
    nz = rng.randn(B, S, 9).astype(np.float32)
    for backend, fused in (("auto", "1"), ("auto", "0"), ("simt", "0")):
        os.environ["FD_VOC_FUSED"] = fused
        try:
            gen = Generator(h, backend=backend).to(dev())
        finally:
            os.environ.pop("FD_VOC_FUSED", None)
        gen.remove_weight_norm()
        gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        wav = gen(T_(mel), T_(f0), rand_ini=T_(ri), sine_noise=T_(nz)).cpu().numpy()
        e = rel_l2(wav, g[f"voc_{name}_wav"])
        print(f"generator[{name}, {backend}, fused={fused}] T=128 rel-L2 vs the reference {e:.2e}")
        assert wav.shape == g[f"voc_{name}_wav"].shape and e < 1e-4


# ------------------------------------------------------------------ utils/audio.py mel
def test_audio_mel_transform_vs_reference(golden):
    g = golden("r2_audio")
    wav = T_(g["au_wav"])
    assert rel_l2(dynamic_range_compression(wav.abs() + 1e-7).cpu().numpy(), g["au_drc"]) < 1e-6
    kw = json.loads(str(g["au_kw_hop256"]))
    for tag, k in (("default", {}), ("hop256", kw)):
        mel = get_mel_transform(**k)(wav).cpu().numpy()
        ref = g[f"au_mel_{tag}"]
        assert mel.shape == ref.shape
        e = rel_l2(mel, ref)
        lm = get_mel_from_audio(wav, **k).cpu().numpy()
        e2 = np.abs(lm - g[f"au_from_audio_{tag}"]).max()
        print(f"get_mel_transform[{tag}] rel-L2 {e:.2e}; get_mel_from_audio max |d log-mel| {e2:.2e}")
        assert e < 5e-5 and e2 < 2e-3


def test_mel_unaligned_hop_matches_aligned_math(golden):
    """speed such that hop*speed is not a multiple of 8 (time-stretch augmentation): gathered-frame path vs float64."""
    from fish_diffusion_b200 import PitchAdjustableMelSpectrogram
    from oracle import mel as omel
    g = golden("mel")
    pam = PitchAdjustableMelSpectrogram()
    spec = pam(T_(g["mel_wav"]), key_shift=0, speed=1.1).cpu().numpy()          # hop 563
    ref = omel.pitch_adjustable_mel(g["mel_wav"].astype(np.float64), key_shift=0, speed=1.1)
    assert spec.shape == ref.shape and rel_l2(spec, ref) < 5e-5


# ------------------------------------------------------------------ checkpoints written by the reference classes
def test_checkpoints_written_by_reference_classes(golden, golden_cfg):
    g = golden("r2_ckpt")
    cfg = golden_cfg["WN_SMALL"]
    ck = torch.load(os.path.join(GOLDEN, "ref_ckpt_small.ckpt"), map_location="cpu", weights_only=False)
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **cfg),
                                 mel_channels=cfg["mel_channels"], spec_min=[-5.0], spec_max=[0.0]))
    sd = formats.lightning_state_dict(ck, "model")
    res = diff.load_state_dict({k[len("diffusion."):]: v for k, v in sd.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    diff = diff.to(dev()).eval()
    with torch.no_grad():
        y = diff.denoise_fn(T_(g["ck_x"]), torch.tensor([500], device=dev()), T_(g["ck_c"]))
    assert rel_l2(y.cpu().numpy(), g["ck_y"]) < 2e-5
    ema = formats.lightning_state_dict(ck, "ema_model")
    assert torch.equal(ema["diffusion.denoise_fn.input_projection.conv.weight"] * 2,
                       sd["diffusion.denoise_fn.input_projection.conv.weight"])
    # vocoder: {"generator": state_dict with weight_g / weight_v} + config.json next to it
    voc = NsfHifiGAN(checkpoint_path=os.path.join(GOLDEN, "ref_generator_small.ckpt"),
                     config_file=os.path.join(GOLDEN, "ref_generator_small.json")).to(dev())
    mel, f0 = g["ck_voc_mel"], g["ck_voc_f0"]
    B, T = f0.shape
    rng = np.random.RandomState(int(g["ck_voc_rseed"]))
    ri = rng.rand(B, 9).astype(np.float32)
    nz = rng.randn(B, T * voc.h["hop_size"], 9).astype(np.float32)
    wav = voc.model(T_(mel), T_(f0), rand_ini=T_(ri), sine_noise=T_(nz)).cpu().numpy()
    assert rel_l2(wav, g["ck_voc_wav"]) < 1e-4


# ------------------------------------------------------------------ DiffSinger assembly (a21 / N1)
def test_diffsinger_forward_features_and_train_step_vs_reference(golden, golden_cfg):
    """archs/diffsinger/diffsinger.py DiffSinger.forward_features / forward of the UNMODIFIED reference (r2_diffsinger.npz):
    same state_dict keys, features and masks, and the training loss + encoder gradients through the native backward (the
    gradient w.r.t. the features is what flows back into the projections)."""
    from fish_diffusion_b200 import DiffSinger, pitch_to_scale
    g = golden("r2_diffsinger")
    wn = golden_cfg["WN_SMALL"]
    E, M = wn["d_encoder"], wn["mel_channels"]
    cfg = dict(text_encoder=dict(type="NaiveProjectionEncoder", input_size=24, output_size=E),
               speaker_encoder=dict(type="NaiveProjectionEncoder", input_size=5, output_size=E, use_embedding=True),
               pitch_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=E, preprocessing=pitch_to_scale),
               pitch_shift_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=E, use_neck=True, neck_size=4),
               energy_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=E),
               diffusion=dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **wn), mel_channels=M,
                              noise_loss="smoothed-l1", sampler_interval=10, spec_min=[-5.0], spec_max=[0.0]))
    model = DiffSinger(cfg)
    sd = {k[len("ds_sd_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("ds_sd_")}
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(dev()).train()
    lens = T_(g["ds_lens"])
    Tm = int(g["ds_contents"].shape[1])
    kw = dict(speakers=T_(g["ds_speakers"]), contents=T_(g["ds_contents"]), contents_lens=lens, contents_max_len=Tm,
              mel_lens=lens, mel_max_len=Tm, pitches=T_(g["ds_pitches"]), pitch_shift=T_(g["ds_pitch_shift"]),
              energy=T_(g["ds_energy"]))
    f = model.forward_features(**kw)
    assert rel_l2(f["features"].detach().cpu().numpy(), g["ds_features"]) < 1e-6
    assert np.array_equal(f["x_masks"].cpu().numpy(), g["ds_x_masks"]) and f["cond_masks"] is f["x_masks"]
    out = model.diffusion.train_step(f["features"], T_(g["ds_mel"]), x_masks=f["x_masks"], cond_masks=f["cond_masks"],
                                     t=T_(g["ds_t"]), noise=T_(g["ds_noise"]))
    assert abs(float(out["loss"]) - float(g["ds_loss"])) < 2e-5 * abs(float(g["ds_loss"]))
    out["loss"].backward()
    for name, p in (("text_w", model.text_encoder.projection.weight), ("pitch_w", model.pitch_encoder.projection.weight),
                    ("spk_w", model.speaker_encoder.embedding.weight)):
        e = rel_l2(p.grad.cpu().numpy(), g[f"ds_g_{name}"])
        print(f"DiffSinger encoder gradient {name}: rel-L2 vs the reference autograd {e:.2e}")
        assert e < 2e-4


def test_diffsinger_fused_conditioner_planes(golden, golden_cfg):
    """The feature projections as ONE GEMM writing the sampler's conditioner planes (DiffSinger.conditioner_planes) equal the
    reference's forward_features followed by the masked split, and the sampler gives the same mel from either."""
    from fish_diffusion_b200 import DiffSinger, pitch_to_scale, _native as N
    from gpu_util import planes_to_f64
    g = golden("r2_diffsinger")
    wn = golden_cfg["WN_SMALL"]
    E, M = wn["d_encoder"], wn["mel_channels"]
    cfg = dict(text_encoder=dict(type="NaiveProjectionEncoder", input_size=24, output_size=E),
               speaker_encoder=dict(type="NaiveProjectionEncoder", input_size=5, output_size=E, use_embedding=True),
               pitch_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=E, preprocessing=pitch_to_scale),
               pitch_shift_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=E, use_neck=True, neck_size=4),
               energy_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=E),
               diffusion=dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **wn), mel_channels=M,
                              noise_loss="smoothed-l1", sampler_interval=100, spec_min=[-5.0], spec_max=[0.0]))
    model = DiffSinger(cfg)
    model.load_state_dict({k[len("ds_sd_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("ds_sd_")})
    model = model.to(dev()).eval()
    lens = T_(g["ds_lens"])
    Tm = int(g["ds_contents"].shape[1])
    kw = dict(speakers=T_(g["ds_speakers"]), contents=T_(g["ds_contents"]), contents_lens=lens, contents_max_len=Tm,
              mel_lens=lens, mel_max_len=Tm, pitches=T_(g["ds_pitches"]), pitch_shift=T_(g["ds_pitch_shift"]),
              energy=T_(g["ds_energy"]))
    assert model._fusable()
    f = model.conditioner_planes(**kw)
    want = g["ds_features"].astype(np.float64).copy()
    want[g["ds_x_masks"]] = 0.0                                 # cond_masks masked_fill (wavenet.py:220-221)
    got = planes_to_f64(f["cond_planes"], N.PREC_F16)
    assert rel_l2(got, want) < 2e-6
    with torch.no_grad():
        ff = model.forward_features(**{**kw, "pitches": T_(g["ds_pitches"])})
        mel_a = model.diffusion(ff["features"], x_masks=ff["x_masks"], cond_masks=ff["cond_masks"], sampler_interval=100,
                                noise_predictor="naive", seed=9)
        mel_b = model.synthesize(**{**kw, "pitches": T_(g["ds_pitches"])}, sampler_interval=100, noise_predictor="naive", seed=9)
    assert rel_l2(mel_b.cpu().numpy(), mel_a.cpu().numpy()) < 1e-5


def test_generator_resblock2_vs_reference(golden):
    """`resblock: "2"` (ResBlock2, models.py:119-158: x = x + conv_d(lrelu(x)) for d in (1, 3)) against the reference."""
    g = golden("r2_voc_resblock2")
    h = json.loads(str(g["rb2_cfg"]))
    sd = {k[len("rb2_sd_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("rb2_sd_")}
    mel, f0 = g["rb2_mel"], g["rb2_f0"]
    B, T = f0.shape
    rng = np.random.RandomState(int(g["rb2_rseed"]))
    ri = rng.rand(B, 9).astype(np.float32)
    nz = rng.randn(B, T * int(np.prod(h["upsample_rates"])), 9).astype(np.float32)
    for backend in ("auto", "simt"):
        gen = Generator(h, backend=backend).to(dev())
        gen.remove_weight_norm()
        res = gen.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        wav = gen(T_(mel), T_(f0), rand_ini=T_(ri), sine_noise=T_(nz)).cpu().numpy()
        e = rel_l2(wav, g["rb2_wav"])
        print(f"generator[resblock 2, {backend}] rel-L2 vs the reference {e:.2e}")
        assert wav.shape == g["rb2_wav"].shape and e < 1e-4
