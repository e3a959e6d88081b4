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
