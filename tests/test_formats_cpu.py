"""CPU tests of the on-disk formats either side of the path (fish_diffusion_b200/formats.py): Lightning / Diff-SVC
checkpoints and preprocessed .npy samples -> collated batches -> DiffSinger.forward_features."""
import numpy as np
import pytest
import torch

from fish_diffusion_b200 import DIFFUSIONS, DiffSinger, ENCODERS, formats, load_checkpoint

WN = dict(mel_channels=16, d_encoder=32, residual_channels=64, residual_layers=2, use_linear_bias=True, dilation_cycle=2)


def make_diffusion(seed):
    torch.manual_seed(seed)
    d = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **WN), mel_channels=16,
                              sampler_interval=10, spec_min=[-5.0], spec_max=[0.0]))
    with torch.no_grad():
        for p in d.parameters():
            p.normal_()
    return d


def test_lightning_state_dict_parts():
    ck = {"state_dict": {"model.a.w": torch.ones(1), "ema_model.a.w": torch.zeros(1), "vocoder.g.w": torch.ones(2)},
          "epoch": 3}
    assert list(formats.lightning_state_dict(ck)) == ["a.w"] and float(formats.lightning_state_dict(ck)["a.w"]) == 1
    assert float(formats.lightning_state_dict(ck, "ema_model")["a.w"]) == 0
    assert list(formats.lightning_state_dict(ck, "vocoder")) == ["g.w"]
    bare = {"a.w": torch.ones(1)}
    assert formats.lightning_state_dict(bare) == bare
    with pytest.raises(KeyError):
        formats.lightning_state_dict({"model.a": torch.ones(1)}, "ema_model")


def test_load_checkpoint_model_and_ema():
    src, dst = make_diffusion(1), make_diffusion(2)
    ema = make_diffusion(3)
    ck = {"state_dict": {**{"model." + k: v for k, v in src.state_dict().items()},
                         **{"ema_model." + k: v for k, v in ema.state_dict().items()},
                         "vocoder.model.conv_pre.weight": torch.zeros(3)}}
    missing, unexpected = load_checkpoint(dst, ck, device="cpu")
    assert not missing and not unexpected
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    load_checkpoint(dst, ck, device="cpu", use_ema=True)
    for k, v in ema.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k


def test_convert_diff_svc_roundtrip_and_checks():
    src, dst = make_diffusion(4), make_diffusion(5)
    enc = lambda n: ENCODERS.build(dict(type="NaiveProjectionEncoder", input_size=n, output_size=32, use_embedding=True))
    pitch, spk = enc(300), enc(4)
    old = {formats.diff_svc_key(k): v.clone() for k, v in src.state_dict().items() if "_noise_predictor" not in k}
    assert "model.denoise_fn.residual_layers.0.dilated_conv.weight" in old       # the Diff-SVC spelling
    assert "model.denoise_fn.mlp.0.weight" in old and "model.denoise_fn.input_projection.bias" in old
    old["model.fs2.pitch_embed.weight"] = torch.randn(300, 32)
    old["model.fs2.spk_embed_proj.weight"] = torch.randn(4, 32)
    old["model.fs2.encoder.layer.weight"] = torch.randn(2, 2)                   # unmapped FastSpeech2 remains
    old["model.K_step"] = torch.tensor(1000)                                    # loose wrapper attribute
    rep = formats.convert_diff_svc({"state_dict": dict(old)}, dst, pitch, spk)
    assert rep["left"] == ["model.fs2.encoder.layer.weight"]
    for k, v in src.state_dict().items():
        if "_noise_predictor" not in k:
            assert torch.equal(dst.state_dict()[k], v), k
    assert torch.equal(pitch.embedding.weight, old["model.fs2.pitch_embed.weight"])
    assert torch.equal(spk.embedding.weight, old["model.fs2.spk_embed_proj.weight"])
    # speaker table absent -> zeroed; wrong residual width / stray keys -> errors naming the config field
    del old["model.fs2.spk_embed_proj.weight"]
    formats.convert_diff_svc({"state_dict": dict(old)}, dst, pitch, spk)
    assert float(spk.embedding.weight.abs().sum()) == 0.0
    bad = dict(old)
    bad["model.denoise_fn.input_projection.weight"] = torch.zeros(128, 16, 1)
    with pytest.raises(ValueError, match="residual_channels"):
        formats.convert_diff_svc(bad, dst)
    bad = dict(old)
    bad["something.else"] = torch.zeros(1)
    with pytest.raises(KeyError, match="not mapped"):
        formats.convert_diff_svc(bad, dst)


def _sample(rng, T, M=16, E=32, energy=False):
    s = dict(path=f"clip{T}.wav", mel=rng.randn(M, T).astype(np.float32), contents=rng.randn(E, T).astype(np.float32),
             pitches=(100 + 50 * rng.rand(T)).astype(np.float32), key_shift=float(rng.randint(-3, 4)), time_stretch=1.0,
             junk="not picked")
    if energy:
        s["energy"] = rng.rand(T).astype(np.float32)
    return s


def test_sample_file_collate_and_feature_projection(tmp_path):
    rng = np.random.RandomState(0)
    lens = [37, 50, 12]
    raws = [_sample(rng, T) for T in lens]
    paths = []
    for i, s in enumerate(raws):
        p = tmp_path / f"{i}.npy"
        np.save(p, s, allow_pickle=True)
        paths.append(p)
    items = [formats.load_sample(p, speaker_id=2) for p in paths] + [None]       # an unreadable file is dropped
    assert items[0]["mel"].shape == (37, 16) and items[0]["contents"].shape == (37, 32) and "junk" not in items[0]
    b = formats.collate_svc(items)
    assert b["mel"].shape == (3, 50, 16) and b["contents"].shape == (3, 50, 32) and b["pitches"].shape == (3, 50, 1)
    assert b["mel_lens"].tolist() == lens and int(b["mel_max_len"]) == 50 and b["contents_lens"].tolist() == lens
    assert b["key_shift"].shape == (3, 1) and b["time_stretch"].shape == (3, 1)
    assert b["speaker"].dtype == torch.int64 and b["speaker"].tolist() == [2, 2, 2]
    for i, (s, T) in enumerate(zip(raws, lens)):
        assert np.array_equal(b["mel"][i, :T].numpy(), s["mel"].T) and float(b["mel"][i, T:].abs().sum()) == 0.0
        assert np.array_equal(b["pitches"][i, :T, 0].numpy(), s["pitches"]) and float(b["pitches"][i, T:].abs().sum()) == 0
    with pytest.raises(ValueError):
        formats.collate_svc([None])
    # energy variant (NaiveSVCPowerDataset)
    e = formats.collate_svc([formats.sample_to_item(dict(_sample(rng, T, energy=True), speaker=0),
                                                    keys=formats.SVC_KEYS + ("energy",)) for T in (5, 9)])
    assert e["energy"].shape == (2, 9, 1)

    # collated batch -> DiffSinger.forward_features (torch modules; the diffusion itself needs the GPU)
    model = DiffSinger(dict(
        text_encoder=dict(type="NaiveProjectionEncoder", input_size=32, output_size=32),
        speaker_encoder=dict(type="NaiveProjectionEncoder", input_size=4, output_size=32, use_embedding=True),
        pitch_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=32),
        pitch_shift_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=32),
        diffusion=dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **WN), mel_channels=16,
                       sampler_interval=10, spec_min=[-5.0], spec_max=[0.0])))
    kw = formats.model_inputs(b)
    assert set(kw) == {"speakers", "contents", "contents_lens", "contents_max_len", "mel", "mel_lens", "mel_max_len",
                       "pitches", "pitch_shift", "phones2mel", "energy"}
    feats = model.forward_features(**{k: v for k, v in kw.items() if k != "mel"})
    assert feats["features"].shape == (3, 50, 32)
    assert feats["x_masks"].shape == (3, 50) and feats["x_masks"][2, 12:].all() and not feats["x_masks"][2, :12].any()


@pytest.mark.parametrize("tag,kw", [("feat", dict(input_size=20)), ("emb", dict(input_size=50, use_embedding_to_input=True))])
def test_fastspeech2_encoder_vs_reference(golden, tag, kw):
    """ENCODERS['FastSpeech2Encoder'] (config #5's text encoder): same state_dict keys / shapes as the
    reference module and the same eval-mode output on a padded batch (golden from the unmodified reference)."""
    import fish_diffusion_b200.fastspeech  # noqa: F401  (registers the encoder)
    g = golden("encoder")
    enc = ENCODERS.build(dict(type="FastSpeech2Encoder", hidden_size=32, num_layers=2, num_heads=2, ffn_kernel_size=9,
                              dropout=0.1, **kw)).eval()
    ref_sd = {k[len(f"enc_{tag}_sd_"):]: g[k] for k in g if k.startswith(f"enc_{tag}_sd_")}
    own = enc.state_dict()
    assert set(own) == set(ref_sd)
    assert all(tuple(own[k].shape) == ref_sd[k].shape for k in own)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in ref_sd.items()})
    mask = torch.from_numpy(g[f"enc_{tag}_mask"])
    with torch.no_grad():
        y = enc(torch.from_numpy(g[f"enc_{tag}_contents"]), mask).numpy()
    ref = g[f"enc_{tag}_y"]
    assert y.shape == ref.shape == (2, 13, 32)
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max(), np.abs(y - ref).max()
    assert np.all(y[1, 8:] == 0)                                   # padded frames are zeroed
    # longer than the 5000-position table: the reversed table is rebuilt for the new length
    enc(torch.zeros(1, 5003, 20) if tag == "feat" else torch.zeros(1, 5003, dtype=torch.long),
        torch.zeros(1, 5003, dtype=torch.bool))
    assert enc._pe.shape[1] == 5003


def test_svs_assembly_features_cpu():
    """BASELINE config #5 (svs_baseline.py:18-26): FastSpeech2Encoder over phoneme ids, gathered to mel frames by
    phones2mel, plus speaker / pitch projections -> features [B, T_mel, 256-like] for the native sampler."""
    torch.manual_seed(0)
    model = DiffSinger(dict(
        text_encoder=dict(type="FastSpeech2Encoder", input_size=40, hidden_size=32, num_layers=2, num_heads=2,
                          use_embedding_to_input=True),
        speaker_encoder=dict(type="NaiveProjectionEncoder", input_size=3, output_size=32, use_embedding=True),
        pitch_encoder=dict(type="NaiveProjectionEncoder", input_size=1, output_size=32),
        diffusion=dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **WN), mel_channels=16,
                       sampler_interval=10, spec_min=[-5.0], spec_max=[0.0]))).eval()
    B, Tph, Tmel = 2, 7, 30
    contents = torch.randint(1, 40, (B, Tph))
    contents_lens = torch.tensor([7, 5])
    mel_lens = torch.tensor([30, 22])
    phones2mel = torch.sort(torch.randint(0, 5, (B, Tmel)), dim=1).values
    pitches = 100 + 50 * torch.rand(B, Tmel, 1)
    with torch.no_grad():
        out = model.forward_features(speakers=torch.tensor([0, 2]), contents=contents, contents_lens=contents_lens,
                                     contents_max_len=Tph, mel_lens=mel_lens, mel_max_len=Tmel, pitches=pitches,
                                     phones2mel=phones2mel)
        enc = model.text_encoder(contents, DiffSinger.get_mask_from_lengths(contents_lens, Tph))
    f = out["features"]
    assert f.shape == (B, Tmel, 32) and out["x_masks"].shape == (B, Tmel)
    # frame t of item b carries phoneme phones2mel[b,t]'s encoding (+ speaker + pitch), padded frames only the additions
    spk = model.speaker_encoder(torch.tensor([0, 2]))[:, None, :]
    pit = model.pitch_encoder(pitches)
    b, t = 1, 10
    want = enc[b, phones2mel[b, t]] + spk[b, 0] + pit[b, t]
    assert torch.allclose(f[b, t], want, atol=1e-5)
    assert torch.allclose(f[1, 25], spk[1, 0] + pit[1, 25], atol=1e-5)


def test_load_pretrained_rules():
    """tools/diffusion/train.py:47-95: EMA weights stand in when the run has no EMA, a speaker table of another size is
    dropped, old predictor-buffer names are tolerated, anything else unexpected is an error."""
    mk = lambda spk: DiffSinger(dict(
        text_encoder=dict(type="NaiveProjectionEncoder", input_size=32, output_size=32),
        speaker_encoder=dict(type="NaiveProjectionEncoder", input_size=spk, output_size=32, use_embedding=True),
        diffusion=dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **WN), mel_channels=16,
                       sampler_interval=10, spec_min=[-5.0], spec_max=[0.0])))
    torch.manual_seed(7)
    src, dst = mk(4), mk(6)
    with torch.no_grad():
        for p in src.parameters():
            p.normal_()
    sd = src.state_dict()
    ck = {"state_dict": {**{"model." + k: torch.zeros_like(v) for k, v in sd.items()},
                         **{"ema_model." + k: v for k, v in sd.items()}, "vocoder.x": torch.zeros(1)}}
    # an old checkpoint: predictor buffers directly under `diffusion.`
    old = {}
    for k, v in ck["state_dict"].items():
        old[k.replace(".naive_noise_predictor.", ".")] = v
    rep = formats.load_pretrained(dst, {"state_dict": old})
    assert rep["dropped"] == ["speaker_encoder.embedding.weight"]
    assert all(".naive_noise_predictor." in k or k == "speaker_encoder.embedding.weight" for k in rep["missing"])
    w = "diffusion.denoise_fn.residual_layers.0.conv_layer.conv.weight"
    assert torch.equal(dst.state_dict()[w], sd[w])                 # the EMA copy, not the zeroed `model.` copy
    assert dst.speaker_encoder.embedding.weight.shape[0] == 6      # own table kept
    # with an EMA in the run the `model.` weights are taken
    formats.load_pretrained(dst, ck, has_ema=True)
    assert float(dst.state_dict()[w].abs().sum()) == 0.0
    bad = dict(ck["state_dict"])
    bad["ema_model.diffusion.denoise_fn.bogus"] = torch.zeros(1)
    with pytest.raises(KeyError, match="unexpected"):
        formats.load_pretrained(dst, {"state_dict": bad})
