"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference, imported by file path from
/root/reference (only possible in the build container -- the GPU box has no /root/reference; the vectors are
committed).  Run:  python tests/golden/make_golden.py

What is loaded from the reference:
  fish_diffusion/modules/wavenet.py                                    (torch only)
  fish_diffusion/archs/diffsinger/diffusions/{uni_pc,noise_predictor,diffusion}.py
        diffusion.py imports `.builder` (mmengine + unrelated denoisers): a stub `.builder` module provides the
        two registries with the reference WaveNet registered, nothing else is touched.
  fish_diffusion/modules/vocoders/nsf_hifigan/models.py                 (numpy + torch)
  fish_diffusion/utils/pitch_adjustable_mel.py   with `librosa.filters.mel` stubbed by the oracle's Slaney
        filterbank (librosa is absent) -> pins padding / STFT / magnitude / key-shift logic, NOT the filterbank.
Every random draw of the reference (torch.randn / randn_like / rand) is served from a recorded numpy stream so the
same numbers can be injected into the CUDA path and the oracle.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import mel as omel  # noqa: E402
from oracle import nsf_hifigan as ovoc  # noqa: E402
from oracle import wavenet as ownet  # noqa: E402
from fish_diffusion_b200.registry import _MiniRegistry  # noqa: E402


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    ref = types.SimpleNamespace()
    ref.wavenet = _load("ref_wavenet", f"{REF}/fish_diffusion/modules/wavenet.py")
    pkg = types.ModuleType("refdiff")
    pkg.__path__ = [f"{REF}/fish_diffusion/archs/diffsinger/diffusions"]
    sys.modules["refdiff"] = pkg
    builder = types.ModuleType("refdiff.builder")
    builder.DIFFUSIONS = _MiniRegistry("diffusions")
    builder.DENOISERS = _MiniRegistry("denoisers")
    builder.DENOISERS.register_module(name="WaveNetDenoiser", module=ref.wavenet.WaveNet)
    sys.modules["refdiff.builder"] = builder
    ref.uni_pc = _load("refdiff.uni_pc", f"{REF}/fish_diffusion/archs/diffsinger/diffusions/uni_pc.py", "refdiff")
    ref.noise_predictor = _load("refdiff.noise_predictor",
                                f"{REF}/fish_diffusion/archs/diffsinger/diffusions/noise_predictor.py", "refdiff")
    ref.diffusion = _load("refdiff.diffusion", f"{REF}/fish_diffusion/archs/diffsinger/diffusions/diffusion.py",
                          "refdiff")
    ref.nsf = _load("ref_nsf_models", f"{REF}/fish_diffusion/modules/vocoders/nsf_hifigan/models.py")
    # pitch_adjustable_mel needs librosa.filters.mel and loguru
    lib = types.ModuleType("librosa")
    filt = types.ModuleType("librosa.filters")
    filt.mel = lambda sr, n_fft, n_mels, fmin, fmax: omel.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    lib.filters = filt
    sys.modules.setdefault("librosa", lib)
    sys.modules.setdefault("librosa.filters", filt)
    ref.mel = _load("ref_pam", f"{REF}/fish_diffusion/utils/pitch_adjustable_mel.py")
    return ref


class RecordedRandom:
    """Serve torch.randn / randn_like / rand from a numpy stream and record every draw."""

    def __init__(self, seed):
        self.rng = np.random.RandomState(seed)
        self.log = []

    def __enter__(self):
        self._o = (torch.randn, torch.randn_like, torch.rand)

        def randn(*size, **kw):
            shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
            a = self.rng.randn(*shape).astype(np.float32)
            self.log.append(("randn", a))
            return torch.from_numpy(a.copy())

        def randn_like(t, **kw):
            a = self.rng.randn(*t.shape).astype(np.float32)
            self.log.append(("randn_like", a))
            return torch.from_numpy(a.copy()).to(t.dtype)

        def rand(*size, **kw):
            shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
            a = self.rng.rand(*shape).astype(np.float32)
            self.log.append(("rand", a))
            return torch.from_numpy(a.copy())

        torch.randn, torch.randn_like, torch.rand = randn, randn_like, rand
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like, torch.rand = self._o


def ref_wavenet(ref, sd, cfg):
    net = ref.wavenet.WaveNet(**cfg)
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.eval()


WN_SMALL = dict(mel_channels=16, d_encoder=32, residual_channels=64, residual_layers=4, use_linear_bias=True,
                dilation_cycle=2)
WN_TC = dict(mel_channels=64, d_encoder=64, residual_channels=128, residual_layers=3, use_linear_bias=True,
             dilation_cycle=4)
WN_NOBIAS = dict(mel_channels=16, d_encoder=32, residual_channels=64, residual_layers=2, use_linear_bias=False,
                 dilation_cycle=None)
WN_FULL = dict(mel_channels=128, d_encoder=256, residual_channels=512, residual_layers=20, use_linear_bias=True,
               dilation_cycle=4)


def wn_weights(seed, cfg):
    c = {k: v for k, v in cfg.items() if k != "dilation_cycle"}
    return ownet.make_wavenet_weights(seed, **c)


def gold_schedules(ref, out):
    for mode in ("linear", "cosine"):
        betas = ref.diffusion.get_noise_schedule_list(mode, 1000, 0.01 if mode == "linear" else 0.02, 0.008)
        out[f"sched_{mode}_betas_f64"] = betas
        naive = ref.noise_predictor.NaiveNoisePredictor(betas=betas)
        for k, v in naive.state_dict().items():
            out[f"sched_{mode}_naive_{k}"] = v.numpy()
        plms = ref.noise_predictor.PLMSNoisePredictor(betas=betas)
        out[f"sched_{mode}_plms_alphas_cumprod"] = plms.alphas_cumprod.numpy()
        ns = ref.uni_pc.NoiseScheduleVP(schedule="discrete", betas=torch.from_numpy(betas))
        out[f"sched_{mode}_unipc_t_array"] = ns.t_array.numpy()
        out[f"sched_{mode}_unipc_log_alpha_array"] = ns.log_alpha_array.numpy()
    # the chunks / t_prev index sequences (integer index math, bit exact)
    for interval, skip in ((1, 0), (5, 970), (10, 0), (100, 0), (10, 970)):
        chunks = torch.arange(0, 1000 - skip, interval, dtype=torch.long).flip(0)
        out[f"chunks_i{interval}_s{skip}"] = chunks.numpy()
        tp = chunks - interval
        out[f"tprev_i{interval}_s{skip}"] = (tp * (tp > 0)).numpy()


def gold_wavenet(ref, out):
    cases = [("small", WN_SMALL, 11, 2, 50), ("tc", WN_TC, 12, 2, 200), ("nobias", WN_NOBIAS, 13, 1, 37),
             ("full", WN_FULL, 0, 2, 128)]
    for name, cfg, seed, B, T in cases:
        sd = wn_weights(seed, cfg)
        net = ref_wavenet(ref, sd, cfg)
        rng = np.random.RandomState(seed + 100)
        x = rng.randn(B, cfg["mel_channels"], T).astype(np.float32)
        cond = rng.randn(B, cfg["d_encoder"], T).astype(np.float32)
        out[f"wn_{name}_x"] = x
        out[f"wn_{name}_cond"] = cond
        with torch.no_grad():
            y = net(torch.from_numpy(x), torch.tensor([990], dtype=torch.long), torch.from_numpy(cond))
            out[f"wn_{name}_y_t990"] = y.numpy()
            if name in ("small", "tc"):
                steps = torch.tensor([17.0, 503.25][:B], dtype=torch.float32)
                out[f"wn_{name}_y_stepsB"] = net(torch.from_numpy(x), steps, torch.from_numpy(cond)).numpy()
                masks = np.zeros((B, T), dtype=bool)
                masks[0, T - 7:] = True
                masks[1, T - 19:] = True
                out[f"wn_{name}_masks"] = masks
                ym = net(torch.from_numpy(x), torch.tensor([40], dtype=torch.long), torch.from_numpy(cond),
                         x_masks=torch.from_numpy(masks), cond_masks=torch.from_numpy(masks))
                out[f"wn_{name}_y_masked_t40"] = ym.numpy()
                y4 = net(torch.from_numpy(x)[:, None], torch.tensor([990], dtype=torch.long), torch.from_numpy(cond))
                assert y4.shape == (B, 1, cfg["mel_channels"], T)
            # fp64 arbiter of the same module
            net64 = ref_wavenet(ref, sd, cfg).double()
            y64 = net64(torch.from_numpy(x).double(), torch.tensor([990.0], dtype=torch.float64),
                        torch.from_numpy(cond).double())
            out[f"wn_{name}_y_t990_f64"] = y64.numpy()


def gold_sampler(ref, out):
    cfg = WN_SMALL
    sd = wn_weights(21, cfg)
    B, T, M, E = 2, 40, cfg["mel_channels"], cfg["d_encoder"]
    rng = np.random.RandomState(22)
    feats = rng.randn(B, T, E).astype(np.float32)
    out["samp_features"] = feats
    for pred in ("naive", "plms", "unipc"):
        for interval, skip in ((100, 0), (50, 900)):
            if skip and pred == "unipc":
                continue
            diff = ref.diffusion.GaussianDiffusion(
                denoiser=dict(type="WaveNetDenoiser", **cfg), mel_channels=M, noise_schedule="linear", timesteps=1000,
                max_beta=0.01, noise_loss="smoothed-l1", sampler_interval=interval, spec_min=[-5.0], spec_max=[0.0],
                noise_predictor=pred)
            diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
            diff.eval()
            orig = None
            if skip:
                orig = (rng.rand(B, T, M).astype(np.float32) * 5 - 5)
                out[f"samp_{pred}_i{interval}_s{skip}_original_mel"] = orig
            with RecordedRandom(1000 + interval) as rr, torch.no_grad():
                y = diff(torch.from_numpy(feats), sampler_interval=interval, skip_steps=skip,
                         original_mel=None if orig is None else torch.from_numpy(orig).transpose(1, 2),
                         noise_predictor=pred)
            key = f"samp_{pred}_i{interval}_s{skip}"
            out[key + "_mel"] = y.numpy()
            for j, (kind, a) in enumerate(rr.log):
                out[key + f"_noise{j}"] = a
            out[key + "_nnoise"] = np.array(len(rr.log))
    # train_step pieces: q_sample + loss with injected t / noise
    diff = ref.diffusion.GaussianDiffusion(
        denoiser=dict(type="WaveNetDenoiser", **cfg), mel_channels=M, noise_loss="smoothed-l1", sampler_interval=10,
        spec_min=[-5.0], spec_max=[0.0])
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    mel = (rng.rand(B, T, M).astype(np.float32) * 5 - 5)
    t = torch.tensor([3, 871], dtype=torch.long)
    noise = rng.randn(B, M, T).astype(np.float32)
    with torch.no_grad():
        x = diff.norm_spec(torch.from_numpy(mel)).transpose(1, 2)
        noised, eps, loss = diff.p_losses(x, t, torch.from_numpy(feats).transpose(1, 2), noise=torch.from_numpy(noise))
    out["train_mel"], out["train_t"], out["train_noise"] = mel, t.numpy(), noise
    out["train_noised"], out["train_eps"], out["train_loss"] = noised.numpy(), eps.numpy(), loss.numpy()


VOC_SMALL = dict(resblock="1", upsample_rates=[4, 4, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4],
                 upsample_initial_channel=128, resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=32, hop_size=64,
                 sampling_rate=44100)


def f0_contour(rng, B, T):
    f0 = 220.0 * 2 ** (rng.randn(B, 1) * 0.5 + 0.3 * np.sin(np.arange(T)[None] / 7.0 + rng.rand(B, 1) * 6))
    f0 = f0.astype(np.float32)
    uv = rng.rand(B, T) < 0.25
    f0[uv] = 0.0
    return f0


def gold_vocoder(ref, out):
    h = ref.nsf.AttrDict(VOC_SMALL)
    sd = ovoc.make_generator_weights(31, VOC_SMALL)
    gen = ref.nsf.Generator(h)
    gen.remove_weight_norm()
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    gen.eval()
    rng = np.random.RandomState(32)
    B, T = 2, 24
    mel = (rng.randn(B, h.num_mels, T) - 2.5).clip(-11.5, 2).astype(np.float32)
    f0 = f0_contour(rng, B, T)
    out["voc_small_mel"], out["voc_small_f0"] = mel, f0
    with RecordedRandom(33) as rr, torch.no_grad():
        wav = gen(torch.from_numpy(mel), torch.from_numpy(f0))
    out["voc_small_wav"] = wav.numpy()
    kinds = [k for k, _ in rr.log]
    assert kinds == ["rand", "randn_like", "randn_like"], kinds
    out["voc_small_rand_ini_raw"] = rr.log[0][1]      # before rand_ini[:,0] = 0
    out["voc_small_sine_noise"] = rr.log[1][1]        # [B,S,9]
    # source module alone, longer, incl. fully voiced / fully unvoiced items
    S_T = 64
    f0b = f0_contour(rng, 3, S_T)
    f0b[1] = 441.0
    f0b[2] = 0.0
    out["src_f0"] = f0b
    with RecordedRandom(34) as rr, torch.no_grad():
        f0_up = torch.nn.functional.interpolate(torch.from_numpy(f0b)[:, None], size=S_T * h.hop_size, mode="linear")
        har, _, _ = gen.m_source(f0_up.transpose(1, 2))
    out["src_f0_up"] = f0_up[:, 0].numpy()
    out["src_har"] = har[:, :, 0].numpy()
    out["src_rand_ini_raw"] = rr.log[0][1]
    out["src_sine_noise"] = rr.log[1][1]
    # weight-norm folding (ckpt format)
    gen2 = ref.nsf.Generator(h)
    sd2 = {k: v.detach().numpy().copy() for k, v in gen2.state_dict().items()}
    gen2.remove_weight_norm()
    sd2f = {k: v.detach().numpy().copy() for k, v in gen2.state_dict().items()}
    for k in ("conv_pre", "ups.0", "resblocks.0.convs1.1", "conv_post"):
        g = sd2.get(k + ".weight_g", sd2.get(k + ".parametrizations.weight.original0"))
        v = sd2.get(k + ".weight_v", sd2.get(k + ".parametrizations.weight.original1"))
        out[f"wnorm_{k}_g"], out[f"wnorm_{k}_v"], out[f"wnorm_{k}_w"] = g, v, sd2f[k + ".weight"]


def gold_mel(ref, out):
    rng = np.random.RandomState(41)
    N = 44100 // 4
    wav = (rng.randn(1, N) * 0.1).astype(np.float32)
    tt = np.arange(N) / 44100.0
    wav[0] += 0.3 * np.sin(2 * np.pi * 330.0 * tt).astype(np.float32)
    out["mel_wav"] = wav
    pam = ref.mel.PitchAdjustableMelSpectrogram(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512,
                                                f_min=40, f_max=16000, n_mels=128)
    for ks in (0, 5, -5):
        out[f"mel_spec_ks{ks}"] = pam(torch.from_numpy(wav), key_shift=ks).numpy()
    out["mel_spec_speed"] = pam(torch.from_numpy(wav), key_shift=0, speed=0.5).numpy()
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(n_freqs=1025, f_min=40.0, f_max=16000.0, n_mels=128, sample_rate=44100,
                                               norm="slaney", mel_scale="slaney").T.numpy()
    out["mel_fb_torchaudio"] = fb.astype(np.float32)


def gold_train(ref, out):
    """Gradients of GaussianDiffusion.p_losses (smoothed-l1) through the reference WaveNet via torch autograd:
    every denoiser parameter and the conditioner, for per-item steps t[B] (training) -- the pin for the native
    backward kernels."""
    for name, cfg, seed, B, T in (("small", WN_SMALL, 51, 2, 40), ("tc", WN_TC, 52, 2, 200)):
        M, E = cfg["mel_channels"], cfg["d_encoder"]
        sd = wn_weights(seed, cfg)
        diff = ref.diffusion.GaussianDiffusion(
            denoiser=dict(type="WaveNetDenoiser", **cfg), mel_channels=M, noise_loss="smoothed-l1", sampler_interval=10,
            spec_min=[-5.0], spec_max=[0.0])
        diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        rng = np.random.RandomState(seed + 1)
        feats = torch.from_numpy(rng.randn(B, T, E).astype(np.float32)).requires_grad_(True)
        mel = (rng.rand(B, T, M).astype(np.float32) * 5 - 5)
        t = torch.tensor([7, 642][:B], dtype=torch.long)
        noise = rng.randn(B, M, T).astype(np.float32)
        x = diff.norm_spec(torch.from_numpy(mel)).transpose(1, 2)
        noised, eps, loss = diff.p_losses(x, t, feats.transpose(1, 2), noise=torch.from_numpy(noise))
        loss.backward()
        out[f"train_{name}_features"], out[f"train_{name}_mel"] = feats.detach().numpy(), mel
        out[f"train_{name}_t"], out[f"train_{name}_noise"] = t.numpy(), noise
        out[f"train_{name}_loss"] = loss.detach().numpy()
        out[f"train_{name}_eps"] = eps.detach().numpy()
        out[f"train_{name}_gfeatures"] = feats.grad.numpy()
        for k, p in diff.denoise_fn.named_parameters():
            out[f"train_{name}_g_{k}"] = p.grad.numpy()


def gold_encoder(ref, out):
    """FastSpeech2Encoder (modules/encoders/fast_speech.py) with a stub `.builder`: eval-mode outputs for a float-feature
    and an embedding-input instance, padded batch."""
    pkg = types.ModuleType("refenc")
    pkg.__path__ = [f"{REF}/fish_diffusion/modules/encoders"]
    sys.modules["refenc"] = pkg
    builder = types.ModuleType("refenc.builder")
    builder.ENCODERS = _MiniRegistry("encoders")
    sys.modules["refenc.builder"] = builder
    fs = _load("refenc.fast_speech", f"{REF}/fish_diffusion/modules/encoders/fast_speech.py", "refenc")
    res = {}
    for tag, kw in (("feat", dict(input_size=20)), ("emb", dict(input_size=50, use_embedding_to_input=True))):
        torch.manual_seed(31)
        enc = fs.FastSpeech2Encoder(hidden_size=32, num_layers=2, num_heads=2, ffn_kernel_size=9, dropout=0.1, **kw).eval()
        with torch.no_grad():
            for prm in enc.parameters():
                prm.normal_(0, 0.3)
        g = torch.Generator().manual_seed(32)
        contents = torch.randn(2, 13, 20, generator=g) if tag == "feat" else torch.randint(0, 50, (2, 13), generator=g)
        mask = torch.arange(13)[None, :] >= torch.tensor([13, 8])[:, None]
        with torch.no_grad():
            y = enc(contents, mask)
        res[f"enc_{tag}_contents"] = contents.numpy()
        res[f"enc_{tag}_mask"] = mask.numpy()
        res[f"enc_{tag}_y"] = y.numpy()
        for k, v in enc.state_dict().items():
            res[f"enc_{tag}_sd_{k}"] = v.numpy()
    out.update(res)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = load_reference()
    groups = {"encoder": gold_encoder,
              "schedules": gold_schedules, "wavenet": gold_wavenet, "sampler": gold_sampler, "vocoder": gold_vocoder,
              "mel": gold_mel, "train": gold_train}
    only = sys.argv[1:]
    for name, fn in groups.items():
        if only and name not in only:
            continue
        out = {}
        fn(ref, out)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")
    # state_dict key / shape inventories of the reference classes (drop-in boundary)
    keys = {}
    net = ref.wavenet.WaveNet(**WN_FULL)
    keys["wavenet_full"] = {k: list(v.shape) for k, v in net.state_dict().items()}
    net = ref.wavenet.WaveNet(**WN_NOBIAS)
    keys["wavenet_nobias"] = {k: list(v.shape) for k, v in net.state_dict().items()}
    diff = ref.diffusion.GaussianDiffusion(denoiser=dict(type="WaveNetDenoiser", **WN_SMALL), mel_channels=16,
                                           spec_min=[-5.0], spec_max=[0.0])
    keys["diffusion_small"] = {k: list(v.shape) for k, v in diff.state_dict().items()}
    for name in ("config_v1", "config_v1_256"):
        with open(f"{REF}/tools/nsf_hifigan/{name}.json") as f:
            hh = ref.nsf.AttrDict(json.load(f))
        g = ref.nsf.Generator(hh)
        keys[f"generator_{name}_wn"] = {k: list(v.shape) for k, v in g.state_dict().items()}
        g.remove_weight_norm()
        keys[f"generator_{name}"] = {k: list(v.shape) for k, v in g.state_dict().items()}
    with open(os.path.join(HERE, "configs.json"), "w") as f:
        json.dump(dict(WN_SMALL=WN_SMALL, WN_TC=WN_TC, WN_NOBIAS=WN_NOBIAS, WN_FULL=WN_FULL, VOC_SMALL=VOC_SMALL,
                       state_dict_keys=keys), f, indent=1)


if __name__ == "__main__":
    main()
