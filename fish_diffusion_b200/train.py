"""Denoiser training step under data parallelism (BASELINE config #4).

Mirrors what the reference does with Lightning (not importable in this image): `DiffSingerLightning` +
`configs/_base_/trainers/base.py:8-41` (DDPStrategy over NCCL, `gradient_as_bucket_view=True`, `static_graph=True`,
`ddp_comm_hook=default_hooks.fp16_compress_hook`, gradient clipping 0.5) and `configs/_base_/schedulers/warmup_cosine.py:13-19`
(AdamW lr 8e-4, weight_decay 1e-2, betas (0.9, 0.98), eps 1e-9).  The only collective on the whole path is the bucketed
gradient all-reduce of torch DDP over NCCL / NVLink; forward and backward are the native kernels (WaveNetTrainFn).
"""
from __future__ import annotations

import torch
from torch import nn


class TrainStepModule(nn.Module):
    """forward(features, mel) -> loss, so that DistributedDataParallel's forward/backward hooks see the step."""

    def __init__(self, diffusion):
        super().__init__()
        self.diffusion = diffusion

    def forward(self, features, mel, t=None, noise=None):
        return self.diffusion.train_step(features, mel, t=t, noise=noise)["loss"]


class DenoiserTrainer:
    def __init__(self, diffusion, lr=8e-4, weight_decay=1e-2, betas=(0.9, 0.98), eps=1e-9, clip=0.5,
                 fp16_compress=False, device=None):
        """fp16_compress=True reproduces the reference's `ddp_comm_hook=fp16_compress_hook` (trainers/base.py:40).  It
        is off by default here: over NVLink 5 the fp32 all-reduce of the 55 M gradients costs ~1.6 ms per step while
        the hook's cast/divide passes cost ~10 ms (measured on 2xB200, profiles/r01_summary.md)."""
        self.diffusion = diffusion
        self.module = TrainStepModule(diffusion)
        self.ddp = None
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            from torch.nn.parallel import DistributedDataParallel as DDP
            dev_ids = None if device is None or device.type != "cuda" else [device.index]
            self.ddp = DDP(self.module, device_ids=dev_ids, gradient_as_bucket_view=True, static_graph=True)
            if fp16_compress:
                self.ddp.register_comm_hook(None, default_hooks.fp16_compress_hook)
        params = list(diffusion.parameters())
        # same update rule as the reference's torch.optim.AdamW; the fused (single multi-tensor kernel) implementation
        fused = all(p.is_cuda for p in params)
        self.opt = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps, fused=fused)
        self.clip = clip

    def step(self, features, mel, t=None, noise=None):
        """One optimisation step; returns the (local) loss tensor."""
        self.opt.zero_grad(set_to_none=True)
        model = self.ddp if self.ddp is not None else self.module
        loss = model(features, mel, t=t, noise=noise)
        loss.backward()
        if self.clip:
            torch.nn.utils.clip_grad_norm_(self.diffusion.parameters(), self.clip)
        self.opt.step()
        return loss.detach()
