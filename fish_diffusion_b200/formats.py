"""On-disk formats either side of the hot path: what the reference stores and how it
reaches the native modules.  Host-side Python only; nothing here touches the GPU.

* Lightning checkpoints of `DiffSingerLightning` (`utils/inference.py:6-32`, `archs/diffsinger/diffsinger.py:182-206`):
  `{"state_dict": {"model.<...>", "ema_model.<...>", "vocoder.<...>"}}`.
* Diff-SVC checkpoints (`tools/diffusion/diff_svc_converter.py:10-116`): the same WaveNet under older key names.
* Preprocessed samples (`tools/preprocessing/extract_features.py:110-172`, `datasets/naive.py:31-85`): one pickled dict per
  `.npy` file with `mel [M,T]`, `contents [E,T]`, `pitches [T]`, `key_shift`, `time_stretch`, optionally `energy [T]`.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------- checkpoints
def lightning_state_dict(ckpt: Mapping, part: str = "model") -> Dict[str, torch.Tensor]:
    """Sub-module weights out of a Lightning checkpoint (or a bare state_dict).

    part = "model" (training weights), "ema_model" (the EMA copy the reference keeps when `ema` is configured,
    diffsinger.py:196-206) or "vocoder".  Keys come back without the prefix.  A dict with no prefixed key at all is
    returned unchanged (already a module state_dict)."""
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    prefixes = ("model.", "ema_model.", "vocoder.")
    if not any(k.startswith(prefixes) for k in sd):
        return dict(sd)
    pre = part + "."
    out = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    if not out:
        raise KeyError(f"checkpoint holds no '{pre}*' keys (found prefixes: "
                       f"{sorted({k.split('.')[0] for k in sd})})")
    return out


def load_pretrained(model: torch.nn.Module, ckpt: Mapping, has_ema: bool = False) -> Dict[str, List[str]]:
    """Warm start from a Lightning checkpoint with the rules of `tools/diffusion/train.py:47-95` (`--pretrained`),
    for `model` = the `DiffSinger` module (Lightning's `.model`):
      * `vocoder.*` keys are dropped;
      * a checkpoint that carries `ema_model.*` while the run keeps no EMA (`has_ema=False`) contributes its EMA weights
        as the model weights;
      * a speaker-embedding table of another size is dropped (rebuilt from scratch), not an error;
      * loading is non-strict, but every unexpected key must be explained by the predictor-buffer drift of old
        checkpoints (buffers stored as `diffusion.<name>` instead of `diffusion.naive_noise_predictor.<name>`).
    Returns {"missing": [...], "unexpected": [...], "dropped": [...]}."""
    sd = dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
    sd = {k: v for k, v in sd.items() if not k.startswith("vocoder.")}
    if not has_ema and any(k.startswith("ema_model.") for k in sd):
        sd = {"model." + k[len("ema_model."):]: v for k, v in sd.items() if k.startswith("ema_model.")}
    sd = lightning_state_dict(sd, "model")
    dropped = []
    spk = "speaker_encoder.embedding.weight"
    own = model.state_dict()
    if spk in sd and spk in own and sd[spk].shape != own[spk].shape:
        del sd[spk]
        dropped.append(spk)
    res = model.load_state_dict(sd, strict=False)
    explained = {k.replace(".naive_noise_predictor.", ".") for k in res.missing_keys}
    stray = sorted(set(res.unexpected_keys) - explained)
    if stray:
        raise KeyError(f"unexpected keys in the checkpoint: {stray[:8]}{' ...' if len(stray) > 8 else ''}")
    return {"missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys), "dropped": dropped}


def diff_svc_key(fish_key: str) -> str:
    """Diff-SVC name of a `GaussianDiffusion.state_dict()` key (diff_svc_converter.py:49-56): the ConvNorm / LinearNorm
    wrappers did not exist there and the dilated conv was called `dilated_conv`."""
    return "model." + fish_key.replace(".conv.", ".").replace(".linear.", ".").replace(".conv_layer.", ".dilated_conv.")


def convert_diff_svc(diff_svc_ckpt: Mapping, diffusion: torch.nn.Module, pitch_encoder: Optional[torch.nn.Module] = None,
                     speaker_encoder: Optional[torch.nn.Module] = None) -> Dict[str, List[str]]:
    """Load a Diff-SVC checkpoint into native modules (same checks and mapping as the reference converter).

    diffusion: GaussianDiffusion (its WaveNet must match `residual_channels` and the spec_min/max width);
    pitch_encoder / speaker_encoder: NaiveProjectionEncoder(use_embedding=True) or None to skip.
    Returns {"loaded": [...], "left": [...]} -- keys consumed and `model.fs2.*` keys that have no counterpart."""
    sd = dict(diff_svc_ckpt["state_dict"] if "state_dict" in diff_svc_ckpt else diff_svc_ckpt)
    own = diffusion.state_dict()
    rc_ckpt = sd["model.denoise_fn.input_projection.weight"].shape[0]
    rc_own = own["denoise_fn.input_projection.conv.weight"].shape[0]
    if rc_ckpt != rc_own:
        raise ValueError(f"residual channels mismatch: checkpoint {rc_ckpt} vs model {rc_own} "
                         "(set model.diffusion.denoiser.residual_channels)")
    widths = (sd["model.spec_min"].shape[-1], sd["model.spec_max"].shape[-1], own["spec_min"].shape[-1])
    if not widths[0] == widths[1] == widths[2]:
        raise ValueError(f"spec_min / spec_max width mismatch: checkpoint {widths[0]}, {widths[1]} vs model {widths[2]}")
    mapped, loaded = {}, []
    for k in own:
        if "_noise_predictor" in k:          # predictor buffers are recomputed from the schedule
            continue
        src = diff_svc_key(k)
        if src not in sd:
            raise KeyError(f"Diff-SVC checkpoint lacks {src} (for {k})")
        mapped[k] = sd.pop(src)
        loaded.append(src)
    for k in [k for k in sd if k.startswith("model.") and k.count(".") == 1]:
        sd.pop(k)                            # loose buffers of the old wrapper (model.betas, model.alphas_cumprod, ...)
    stray = [k for k in sd if not k.startswith("model.fs2")]
    if stray:
        raise KeyError(f"keys not mapped: {stray[:8]}{' ...' if len(stray) > 8 else ''}")
    res = diffusion.load_state_dict(mapped, strict=False)
    assert all("_noise_predictor" in k for k in res.missing_keys) and not res.unexpected_keys, res
    if pitch_encoder is not None:
        pitch_encoder.load_state_dict({"embedding.weight": sd.pop("model.fs2.pitch_embed.weight")}, strict=True)
        loaded.append("model.fs2.pitch_embed.weight")
    if speaker_encoder is not None:
        if "model.fs2.spk_embed_proj.weight" in sd:
            w = sd.pop("model.fs2.spk_embed_proj.weight")
            if w.shape[0] != speaker_encoder.embedding.weight.shape[0]:
                raise ValueError(f"speaker count mismatch: checkpoint {w.shape[0]} vs model "
                                 f"{speaker_encoder.embedding.weight.shape[0]} (set speaker_encoder.input_size)")
            speaker_encoder.load_state_dict({"embedding.weight": w}, strict=True)
            loaded.append("model.fs2.spk_embed_proj.weight")
        else:
            with torch.no_grad():
                speaker_encoder.embedding.weight.zero_()
    return {"loaded": loaded, "left": sorted(sd)}


# ---------------------------------------------------------------------------------------------- preprocessed samples
SVC_KEYS = ("path", "time_stretch", "mel", "contents", "pitches", "key_shift", "speaker")


def load_sample(path, speaker_id: int = 0) -> dict:
    """One preprocessed `.npy` sample (a pickled dict) in the per-item layout of NaiveSVCDataset.get_item:
    mel [T,M], contents [T,E] (time-major, i.e. already the channels-last rows the native kernels read), pitches [T]."""
    x = np.load(path, allow_pickle=True).item()
    x["speaker"] = speaker_id
    return sample_to_item(x)


def sample_to_item(x: Mapping, keys: Sequence[str] = SVC_KEYS) -> dict:
    item = {k: x[k] for k in keys if k in x or k not in ("path", "energy")}
    for k in ("mel", "contents"):
        item[k] = np.ascontiguousarray(np.asarray(item[k]).T)          # [C,T] on disk -> [T,C]
    return item


def _pad_stack(arrs: Sequence, axis: int):
    ts = [torch.as_tensor(np.asarray(a)).float() for a in arrs]
    lens = torch.tensor([t.shape[axis] for t in ts], dtype=torch.long)
    L = int(lens.max())
    out = []
    for t in ts:
        pad = [0, 0] * t.dim()
        ax = axis % t.dim()
        pad[2 * (t.dim() - 1 - ax) + 1] = L - t.shape[ax]              # F.pad lists the last dimension first
        out.append(torch.nn.functional.pad(t, pad))
    return torch.stack(out), lens, torch.tensor(L)


def collate_svc(items: Iterable[Optional[dict]]) -> dict:
    """Batch of items -> the dict `DiffSingerLightning._step` consumes (NaiveSVCDataset.collating_pipeline,
    datasets/naive.py:68-85): zero-padded `mel [B,T,M]`, `contents [B,T,E]`, `pitches [B,T,1]` with their `_lens` /
    `_max_len`, `time_stretch [B,1]`, `key_shift [B,1]`, `speaker [B]` (int64); `energy [B,T,1]` when present.
    `None` items (unreadable files) are dropped, as in NaiveDataset.collate_fn."""
    items = [i for i in items if i is not None]
    if not items:
        raise ValueError("empty batch")
    out: dict = {}
    for k, axis in (("mel", -2), ("contents", -2), ("pitches", -1), ("energy", -1)):
        if k in items[0]:
            out[k], out[k + "_lens"], out[k + "_max_len"] = _pad_stack([i[k] for i in items], axis)
    for k, dt in (("time_stretch", torch.float32), ("key_shift", torch.float32), ("speaker", torch.int64)):
        if k in items[0]:
            out[k] = torch.tensor([i[k] for i in items], dtype=dt)
    for k in ("pitches", "energy", "time_stretch", "key_shift"):
        if k in out:
            out[k] = out[k].unsqueeze(-1)
    if "path" in items[0]:
        out["path"] = [i["path"] for i in items]
    return out


def model_inputs(batch: Mapping) -> dict:
    """Keyword arguments of `DiffSinger.forward` / `forward_features` from a collated batch
    (DiffSingerLightning._step, diffsinger.py:259-288)."""
    if batch.get("pitches") is not None and "mel" in batch:
        assert batch["pitches"].shape[1] == batch["mel"].shape[1], "pitches and mel disagree on the frame count"
    return dict(speakers=batch.get("speaker"), contents=batch["contents"], contents_lens=batch["contents_lens"],
                contents_max_len=batch["contents_max_len"], mel=batch.get("mel"), mel_lens=batch["mel_lens"],
                mel_max_len=batch["mel_max_len"], pitches=batch.get("pitches"), pitch_shift=batch.get("key_shift"),
                phones2mel=batch.get("phones2mel"), energy=batch.get("energy"))
