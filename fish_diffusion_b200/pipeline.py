"""Batched inference driver.

The reference synthesises strictly one segment at a time (``tools/diffusion/inference.py:133-160``,
``inference_svs.py:97-149``): features -> 100-step sampler -> vocoder with B = 1.  This driver is the caller the
BASELINE configs assume: variable-length segments are bucketed by length, padded to a common T with the reference's
mask semantics (``x_masks`` / ``cond_masks`` = True on padding, exactly what ``DiffSinger.forward_features`` returns),
sampled as one batch, and vocoded on a second CUDA stream while the sampler of the next batch runs.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def plan_batches(lengths: Sequence[int], max_batch: int = 32, bucket: int = 128, max_frames: int = None,
                 max_waste: float = 0.1, min_batch: int = 4):
    """Group item indices into batches.  Items are sorted by length (longest first), every batch is padded to the
    next multiple of `bucket` of its longest item, holds at most `max_batch` items and at most `max_frames` padded
    frames in total; a batch of at least `min_batch` items is closed early when the next item would be padded by more
    than `max_waste` of the batch length.  Returns [(indices, T_pad), ...]; every index appears exactly once."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    batches, cur, cur_T = [], [], 0
    for i in order:
        L = int(lengths[i])
        if L <= 0:
            raise ValueError(f"item {i} has non-positive length {L}")
        T_pad = (L + bucket - 1) // bucket * bucket
        if not cur:
            cur, cur_T = [i], T_pad
            continue
        full = len(cur) >= max_batch or (max_frames is not None and (len(cur) + 1) * cur_T > max_frames)
        full = full or (len(cur) >= min_batch and T_pad < cur_T * (1.0 - max_waste))
        if full:
            batches.append((cur, cur_T))
            cur, cur_T = [i], T_pad
        else:
            cur.append(i)          # sorted descending: cur_T already covers this item
    if cur:
        batches.append((cur, cur_T))
    return batches


def padding_waste(lengths: Sequence[int], batches) -> float:
    used = sum(int(l) for l in lengths)
    padded = sum(len(idx) * T for idx, T in batches)
    return 1.0 - used / max(padded, 1)


class BatchedSynthesizer:
    """features/f0 segments -> mel (native sampler) -> waveform (native NSF-HiFiGAN), batched."""

    def __init__(self, diffusion, generator, max_batch: int = 32, bucket: int = 128, max_frames: int = None,
                 sampler_interval: int = None, noise_predictor: str = None, hop: int = None):
        self.diffusion, self.generator = diffusion, generator
        self.max_batch, self.bucket, self.max_frames = max_batch, bucket, max_frames
        self.sampler_interval, self.noise_predictor = sampler_interval, noise_predictor
        import numpy as np
        self.hop = hop if hop is not None else int(np.prod(generator.h.upsample_rates))

    @torch.no_grad()
    def __call__(self, features: List[torch.Tensor], f0: List[torch.Tensor], seed: int = 0, return_mel: bool = False):
        """features[i] [T_i, E], f0[i] [T_i] (CUDA tensors) -> list of wav [T_i * hop] (and mel [T_i, M])."""
        assert len(features) == len(f0) and len(features) > 0
        dev = features[0].device
        lengths = [int(x.shape[0]) for x in features]
        batches = plan_batches(lengths, self.max_batch, self.bucket, self.max_frames)
        E = features[0].shape[1]
        wavs, mels = [None] * len(features), [None] * len(features)
        s_main = torch.cuda.current_stream(dev)
        s_voc = torch.cuda.Stream(device=dev)
        pending = []
        for bi, (idx, T) in enumerate(batches):
            B = len(idx)
            feat = torch.zeros((B, T, E), dtype=torch.float32, device=dev)
            pitch = torch.zeros((B, T), dtype=torch.float32, device=dev)
            mask = torch.ones((B, T), dtype=torch.bool, device=dev)
            for j, i in enumerate(idx):
                feat[j, :lengths[i]] = features[i]
                pitch[j, :lengths[i]] = f0[i]
                mask[j, :lengths[i]] = False
            mel = self.diffusion(feat, sampler_interval=self.sampler_interval, noise_predictor=self.noise_predictor,
                                 x_masks=mask, cond_masks=mask, seed=seed + bi)            # [B,T,M]
            ev = torch.cuda.Event()
            ev.record(s_main)
            with torch.cuda.stream(s_voc):
                s_voc.wait_event(ev)
                mel.record_stream(s_voc)
                pitch.record_stream(s_voc)
                wav = self.generator(mel.transpose(1, 2).contiguous(), pitch, seed=seed + bi)[:, 0]   # [B, T*hop]
            pending.append((idx, mel, wav))
        s_main.wait_stream(s_voc)
        for idx, mel, wav in pending:
            for j, i in enumerate(idx):
                wavs[i] = wav[j, :lengths[i] * self.hop]
                mels[i] = mel[j, :lengths[i]]
        return (wavs, mels) if return_mel else wavs
