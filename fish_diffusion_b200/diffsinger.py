"""Feature-projection front of the sampler: the reference's ``DiffSinger`` assembly
(``fish_diffusion/archs/diffsinger/diffsinger.py:20-179``) with ``NaiveProjectionEncoder``
(``fish_diffusion/modules/encoders/naive_projection.py:6-60``) and ``pitch_to_scale`` (``utils/pitch.py:12-22``).

The encoders are tiny Linear / Embedding layers on [B,T,<=256] tensors and stay ordinary torch modules, exactly as
SURVEY.md scopes them; what matters for the hot path is that ``features`` leave this module channels-last
``[B,T,E]`` -- already the layout the native sampler consumes (it is split into planes once per sampler call, never
transposed).  Same module / parameter names as the reference, so ``model.*`` keys of a Lightning checkpoint load.
"""
from __future__ import annotations

import torch
from torch import nn

from .registry import DIFFUSIONS, Registry

ENCODERS = Registry("encoders")

_f0_max = 1100.0
_f0_min = 50.0


def pitch_to_scale(f0, f0_min=_f0_min, f0_max=_f0_max):
    """utils/pitch.py:12-22."""
    f0_scale = (f0 - f0_min) / (f0_max - f0_min)
    f0_scale = f0_scale.clamp(0, 1)
    if f0.ndim == 2:
        f0_scale = f0_scale.unsqueeze(-1)
    return f0_scale


@ENCODERS.register_module(name="NaiveProjectionEncoder", force=True)
class NaiveProjectionEncoder(nn.Module):
    def __init__(self, input_size, output_size, use_embedding: bool = False, use_neck: bool = False, neck_size: int = 8,
                 preprocessing=None):
        super().__init__()
        self.use_embedding, self.input_size, self.output_size = use_embedding, input_size, output_size
        self.preprocessing = preprocessing
        if use_embedding:
            self.embedding = nn.Embedding(input_size, output_size)
        elif use_neck:
            self.projection = nn.Sequential(nn.Linear(input_size, neck_size), nn.Linear(neck_size, output_size))
        else:
            self.projection = nn.Linear(input_size, output_size)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.0)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, mean=0, std=m.embedding_dim ** -0.5)

    def forward(self, x, *args, **kwargs):
        if self.preprocessing is not None:
            x = self.preprocessing(x)
        return self.embedding(x) if self.use_embedding else self.projection(x)


def _cfg_get(cfg, key):
    return cfg.get(key) if isinstance(cfg, dict) else getattr(cfg, key, None)


class DiffSinger(nn.Module):
    """Reference model assembly (diffsinger.py:20-179) on top of the native GaussianDiffusion."""

    def __init__(self, model_config):
        super().__init__()
        self.text_encoder = ENCODERS.build(_cfg_get(model_config, "text_encoder"))
        self.diffusion = DIFFUSIONS.build(_cfg_get(model_config, "diffusion"))
        for name in ("speaker_encoder", "pitch_encoder", "pitch_shift_encoder", "energy_encoder"):
            c = _cfg_get(model_config, name)
            if c:
                setattr(self, name, ENCODERS.build(c))

    @staticmethod
    def get_mask_from_lengths(lengths, max_len=None):
        if max_len is None:
            max_len = int(torch.max(lengths).item())
        ids = torch.arange(0, max_len, device=lengths.device).unsqueeze(0).expand(lengths.shape[0], -1)
        return ids >= lengths.unsqueeze(1).expand(-1, max_len)

    def forward_features(self, speakers, contents, contents_lens, contents_max_len, mel_lens=None, mel_max_len=None,
                         pitches=None, pitch_shift=None, phones2mel=None, energy=None):
        src_masks = self.get_mask_from_lengths(contents_lens, contents_max_len) if contents_lens is not None else None
        mel_masks = self.get_mask_from_lengths(mel_lens, mel_max_len) if mel_lens is not None else None
        features = self.text_encoder(contents, src_masks)
        if phones2mel is not None:
            idx = phones2mel.unsqueeze(-1).repeat([1, 1, features.shape[-1]]).long()
            features = torch.gather(features, 1, idx) * (1 - mel_masks[:, :, None].float())
        if speakers is not None and speakers.ndim in [2, 3] and torch.is_floating_point(speakers):
            speaker_embed = speakers
        elif speakers is not None and hasattr(self, "speaker_encoder"):
            speaker_embed = self.speaker_encoder(speakers)
        else:
            speaker_embed = None
        if speaker_embed is not None and speaker_embed.ndim == 2:
            speaker_embed = speaker_embed[:, None, :]
        if speaker_embed is not None:
            features = features + speaker_embed
        if hasattr(self, "pitch_encoder"):
            features = features + self.pitch_encoder(pitches)
        if pitch_shift is not None and hasattr(self, "pitch_shift_encoder"):
            e = self.pitch_shift_encoder(pitch_shift)
            features = features + (e[:, None, :] if e.ndim == 2 else e)
        if energy is not None and hasattr(self, "energy_encoder"):
            e = self.energy_encoder(energy)
            features = features + (e[:, None, :] if e.ndim == 2 else e)
        return dict(features=features, x_masks=mel_masks, x_lens=mel_lens, cond_masks=mel_masks)

    # ------------------------------------------------------------------------------------ fused front of the sampler
    def _fusable(self):
        """The projection front can be one GEMM when every encoder is a NaiveProjectionEncoder of the plain kind the SVC
        configs use (configs/_base_/archs/diff_svc_v2.py): Linear text encoder, Linear(1 -> E) pitch / energy encoders."""
        te = self.text_encoder
        ok = isinstance(te, NaiveProjectionEncoder) and not te.use_embedding and isinstance(te.projection, nn.Linear) \
            and te.preprocessing is None
        for name in ("pitch_encoder", "energy_encoder"):
            e = getattr(self, name, None)
            ok = ok and (e is None or (isinstance(e, NaiveProjectionEncoder) and not e.use_embedding
                                       and isinstance(e.projection, nn.Linear) and e.input_size == 1))
        return ok

    @torch.no_grad()
    def conditioner_planes(self, speakers, contents, contents_lens, contents_max_len, mel_lens=None, mel_max_len=None,
                           pitches=None, pitch_shift=None, energy=None):
        """`forward_features` (diffsinger.py:57-134) written straight into the sampler's conditioner plane buffer by ONE
        tap-GEMM: K = [contents (E_in) | pitch_scale, energy, 0...], W = [W_text | w_pitch | w_energy | 0], per-item bias =
        b_text + b_pitch + b_energy + speaker embedding + pitch-shift embedding, masked rows zeroed in the epilogue
        (the `cond_masks` masked_fill of WaveNet.forward).  Saves the fp32 [B,T,E] features tensor and its split pass
        (2 x 131 MB at B=32, T=4000).  Returns dict(cond_planes=[2,B,T,E], x_masks, x_lens, cond_masks); falls back to
        forward_features + split for encoder types it does not cover (FastSpeech2 text encoder, necks, embeddings)."""
        from . import _native as N
        if not self._fusable() or not contents.is_cuda:
            f = self.forward_features(speakers, contents, contents_lens, contents_max_len, mel_lens, mel_max_len, pitches,
                                      pitch_shift, None, energy)
            prec = N.prec_code(getattr(self.diffusion.denoise_fn, "precision", "f16"))
            m = None if f["cond_masks"] is None else f["cond_masks"].to(torch.uint8).contiguous()
            f["cond_planes"] = N.split_nwc(f["features"].to(torch.float32), prec, mask=m)
            return f
        dev = contents.device
        den = self.diffusion.denoise_fn
        prec, mma = N.prec_code(den.precision), N.mma_code(den.precision)
        B, T, Ein = contents.shape
        E = self.text_encoder.output_size
        mel_masks = self.get_mask_from_lengths(mel_lens, mel_max_len) if mel_lens is not None else None
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)
        # per-item bias: every term of forward_features that does not depend on t
        bias = f32(self.text_encoder.projection.bias).expand(B, E).clone() if self.text_encoder.projection.bias is not None \
            else torch.zeros((B, E), dtype=torch.float32, device=dev)
        if speakers is not None and speakers.ndim == 2 and torch.is_floating_point(speakers):
            bias += f32(speakers)
        elif speakers is not None and speakers.ndim == 3:
            raise NotImplementedError("per-frame speaker embeddings: use forward_features")
        elif speakers is not None and hasattr(self, "speaker_encoder"):
            se = self.speaker_encoder(speakers)
            if se.ndim != 2:
                raise NotImplementedError("per-frame speaker embeddings: use forward_features")
            bias += f32(se)
        if pitch_shift is not None and hasattr(self, "pitch_shift_encoder"):
            pe = self.pitch_shift_encoder(pitch_shift)
            bias += f32(pe if pe.ndim == 2 else pe[:, 0])
        # per-frame scalar inputs ride in a second source of 64 columns (K block of the tensor-core kernel)
        aux = torch.zeros((B, T, 64), dtype=torch.float32, device=dev)
        w_aux = torch.zeros((E, 64), dtype=torch.float32, device=dev)
        col = 0
        for name, val in (("pitch_encoder", pitches), ("energy_encoder", energy)):
            enc = getattr(self, name, None)
            if enc is None or val is None:
                continue
            v = val if enc.preprocessing is None else enc.preprocessing(val)
            aux[:, :, col] = f32(v).reshape(B, T)
            w_aux[:, col] = f32(enc.projection.weight)[:, 0]
            if enc.projection.bias is not None:
                bias += f32(enc.projection.bias)
            col += 1
        Kp = (Ein + 63) // 64 * 64
        W = torch.zeros((E, Kp + 64), dtype=torch.float32, device=dev)
        W[:, :Ein] = f32(self.text_encoder.projection.weight)
        W[:, Kp:] = w_aux
        s = N.pow2_scale(W)
        wp = N.pack_weight(W, prec, s)
        if Kp != Ein:
            cpad = torch.zeros((B, T, Kp), dtype=torch.float32, device=dev)
            cpad[:, :, :Ein] = contents
            contents = cpad
        src0 = N.split_nwc(f32(contents).contiguous(), prec)
        src1 = N.split_nwc(aux, prec)
        ws = self.diffusion._sampler_ws(dev, B, T, self.diffusion.mel_bins, E)
        m = None if mel_masks is None else mel_masks.to(torch.uint8).contiguous()
        N.gemm_cl(src0, Kp, wp, E, Kp + 64, B, T, [(0, 0, 0, Kp), (1, 0, 0, 64)], src1=src1, C1=64, bias=bias.contiguous(),
                  bias_per_item=True, row_mask=m, out_planes=ws["cond_planes"], w_inv_scale=1.0 / s, prec=mma,
                  backend=N.BACKEND_TC if N.tc_supported_linear(E, 64, 2) else N.BACKEND_SIMT)
        return dict(cond_planes=ws["cond_planes"], x_masks=mel_masks, x_lens=mel_lens, cond_masks=mel_masks)

    @torch.no_grad()
    def synthesize(self, speakers, contents, contents_lens, contents_max_len, mel_lens=None, mel_max_len=None, pitches=None,
                   pitch_shift=None, energy=None, **sampler_kw):
        """features -> mel in two steps on the device: the fused projection GEMM, then the native sampler on its planes
        (what `SVCInference.forward` does through forward_features + diffusion, tools/diffusion/inference.py:85-131)."""
        f = self.conditioner_planes(speakers, contents, contents_lens, contents_max_len, mel_lens, mel_max_len, pitches,
                                    pitch_shift, energy)
        return self.diffusion(None, x_masks=f["x_masks"], cond_planes=f["cond_planes"], **sampler_kw)

    def forward(self, speakers, contents, contents_lens, contents_max_len, mel=None, mel_lens=None, mel_max_len=None,
                pitches=None, pitch_shift=None, phones2mel=None, energy=None):
        features = self.forward_features(speakers=speakers, contents=contents, contents_lens=contents_lens,
                                         contents_max_len=contents_max_len, mel_lens=mel_lens, mel_max_len=mel_max_len,
                                         pitches=pitches, pitch_shift=pitch_shift, phones2mel=phones2mel, energy=energy)
        out = self.diffusion.train_step(features["features"], mel, x_masks=features["x_masks"],
                                        cond_masks=features["cond_masks"])
        out["features"], out["x_masks"] = features["features"], features["x_masks"]
        out["x_lens"], out["cond_masks"] = features["x_lens"], features["cond_masks"]
        return out


def load_checkpoint(model: nn.Module, checkpoint, device="cuda", strict: bool = False, use_ema: bool = False):
    """Reference ``utils/inference.py:6-32`` semantics: Lightning ``state_dict`` with ``model.`` prefixes, ``vocoder.*``
    keys dropped, non-strict.  ``use_ema=True`` takes the ``ema_model.*`` copy instead (the weights the reference
    validates with, diffsinger.py:259-263).  Returns the (missing, unexpected) key lists."""
    from .formats import lightning_state_dict
    state = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, (str, bytes)) else checkpoint
    state = lightning_state_dict(state, "ema_model" if use_ema else "model")
    res = model.load_state_dict(state, strict=strict)
    model.to(device)
    return res.missing_keys, res.unexpected_keys
