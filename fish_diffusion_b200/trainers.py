"""The `DiffSingerLightning` role without Lightning (archs/diffsinger/diffsinger.py:182-405 + configs/_base_/trainers/base.py
+ configs/_base_/schedulers/warmup_cosine.py), for images where pytorch_lightning / mmengine are absent (this one) and as the
explicit statement of what the Lightning module does around the native model:

  model / ema_model / vocoder sub-modules under exactly those names -> `state_dict()` has the `model.*`, `ema_model.*`,
  `vocoder.*` layout of a reference checkpoint (formats.lightning_state_dict reads it back);
  `_step`                  batch dict -> DiffSinger.forward keyword arguments (formats.model_inputs), EMA weights in "valid" mode;
  `training_step`          loss, backward (native WaveNet backward), gradient all-reduce, clip 0.5, AdamW(lr = 1.0) under
                           LambdaLR(warm-up cosine), EMA update with the two `_foreach` calls of the reference;
  `validation_step`        loss with the EMA weights + the sampled mel (and audio when a vocoder is attached).

Everything numerical below the model call is the native path (DiffSinger -> GaussianDiffusion.train_step -> WaveNetTrainFn);
this file is host logic only and is covered by CPU tests (tests/test_trainers_cpu.py).
"""
from __future__ import annotations

import copy
import math
from typing import Callable, Mapping, Optional

import torch
from torch import nn

from .formats import model_inputs


class WarmupCosine:
    """`LambdaWarmUpCosineScheduler` (fish_diffusion/schedulers/warmup_cosine_scheduler.py:6-56): multiplier for a base
    learning rate of 1.0 -- linear from val_start to val_base over warm_up_steps, then half a cosine down to val_final at
    max_decay_steps (held afterwards).  Defaults: configs/_base_/schedulers/warmup_cosine.py:5-11."""

    def __init__(self, *, val_base=8e-4, val_final=2e-5, max_decay_steps=300000, val_start=1e-5, warm_up_steps=1000):
        self.val_base, self.val_final, self.val_start = val_base, val_final, val_start
        self.warm_up_steps, self.max_decay_steps = warm_up_steps, max_decay_steps
        self.last_lr = 0.0

    def __call__(self, n):
        if n < self.warm_up_steps:
            lr = (self.val_base - self.val_start) / self.warm_up_steps * n + self.val_start
        else:
            t = min((n - self.warm_up_steps) / (self.max_decay_steps - self.warm_up_steps), 1.0)
            lr = self.val_final + 0.5 * (self.val_base - self.val_final) * (1 + math.cos(t * torch.pi))
        self.last_lr = lr
        return lr


def ema_update(ema_model: nn.Module, model: nn.Module, momentum: float):
    """diffsinger.py:388-400: ema <- momentum * ema + (1 - momentum) * param over the parameters, as two multi-tensor ops
    (buffers are not averaged, exactly like the reference)."""
    ema_params = [p.data for p in ema_model.parameters()]
    params = [p.data for p in model.parameters()]
    torch._foreach_mul_(ema_params, momentum)
    torch._foreach_add_(ema_params, params, alpha=1.0 - momentum)


class DiffSingerTrainer(nn.Module):
    """model: a fish_diffusion_b200.DiffSinger (or any module whose forward(**model_inputs(batch)) returns a dict with
    "loss", and optionally "features" / "x_masks" / "cond_masks" / "metrics" like the reference's).
    ema_momentum=None disables the EMA copy (config key `ema_momentum`, diffsinger.py:189, 196-206).
    reduce_grads(params): gradient all-reduce hook for one-process-per-GPU data parallelism (the DDPStrategy of
    configs/_base_/trainers/base.py:30-41); `train.GradSync` can additionally be attached to the denoiser for the
    bucketed overlap (see train.DenoiserTrainer)."""

    def __init__(self, model: nn.Module, vocoder: Optional[nn.Module] = None, ema_momentum: Optional[float] = None,
                 optimizer: Optional[Mapping] = None, lr_lambda: Optional[Callable[[int], float]] = None,
                 gradient_clip_val: Optional[float] = 0.5, reduce_grads: Optional[Callable] = None,
                 model_factory: Optional[Callable[[], nn.Module]] = None):
        """model_factory: builds a second, fresh instance for the EMA copy, as the reference does (`model_fn(config.model)`,
        diffsinger.py:197); without it the model is deep-copied, which is only safe before its first use (a native module
        that has already run holds work buffers and captured CUDA graphs)."""
        super().__init__()
        self.model = model
        self.ema_momentum = ema_momentum
        if ema_momentum is not None:
            self.ema_model = model_factory() if model_factory is not None else copy.deepcopy(model)
            self.ema_model.load_state_dict(model.state_dict())       # then the same weights (diffsinger.py:203)
            self.ema_model.eval()
            for p in self.ema_model.parameters():
                p.requires_grad = False
        if vocoder is not None:
            self.vocoder = vocoder
            if hasattr(vocoder, "freeze"):                   # diffsinger.py:212-213
                vocoder.freeze()
            else:
                vocoder.eval()
                for p in vocoder.parameters():
                    p.requires_grad = False
        self.optimizer_cfg = dict(lr=1.0, weight_decay=1e-2, betas=(0.9, 0.98), eps=1e-9)        # warmup_cosine.py:13-19
        if optimizer:
            self.optimizer_cfg.update({k: v for k, v in optimizer.items() if k != "type"})
        self.lr_lambda = lr_lambda if lr_lambda is not None else WarmupCosine()
        self.gradient_clip_val = gradient_clip_val
        self.reduce_grads = reduce_grads
        self.optimizer = self.scheduler = None
        self.logged = {}
        self.global_step = 0

    # ---------------------------------------------------------------------------------------- Lightning-shaped API
    def configure_optimizers(self):
        """diffsinger.py:240-257: every parameter of the module that requires grad (the frozen EMA copy / vocoder hold none)."""
        params = [p for p in self.parameters() if p.requires_grad]
        fused = bool(params) and all(p.is_cuda for p in params)
        self.optimizer = torch.optim.AdamW(params, fused=fused, **self.optimizer_cfg)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda=self.lr_lambda)
        return [self.optimizer], dict(scheduler=self.scheduler, interval="step")

    def log(self, name, value, **kw):
        self.logged[name] = float(value)

    def _step(self, batch, mode):
        """diffsinger.py:259-310 up to the returned loss; in "valid" mode with an EMA copy the EMA weights are used and the
        sampled mel comes back too."""
        model = self.ema_model if (self.ema_momentum is not None and mode == "valid") else self.model
        if "pitches" not in batch:
            batch["pitches"] = None
        output = model(**model_inputs(batch))
        self.log(f"{mode}_loss", output["loss"])
        for k, v in output.get("metrics", {}).items():
            self.log(f"{mode}_{k}", v)
        if mode != "valid":
            return output["loss"], None
        mel = model.diffusion(output["features"], x_masks=output["x_masks"], cond_masks=output["cond_masks"])
        return output["loss"], mel

    def training_step(self, batch):
        if self.optimizer is None:
            self.configure_optimizers()
        self.optimizer.zero_grad(set_to_none=True)
        loss, _ = self._step(batch, "train")
        loss.backward()
        params = [p for p in self.model.parameters() if p.grad is not None]
        if self.reduce_grads is not None:
            self.reduce_grads(params)
        if self.gradient_clip_val:
            torch.nn.utils.clip_grad_norm_(params, self.gradient_clip_val)       # trainers/base.py:11-12 ("norm")
        self.optimizer.step()
        self.scheduler.step()                                                    # interval="step"
        self.global_step += 1
        if self.ema_momentum is not None:
            ema_update(self.ema_model, self.model, self.ema_momentum)
        return loss.detach()

    @torch.no_grad()
    def validation_step(self, batch):
        """-> dict(loss, mel [B,T,M] sampled with the (EMA) weights, wavs: list of per-item audio when a vocoder with
        `spec2wav` is attached and the batch carries pitches -- the tensors viz_synth_sample renders, diffsinger.py:312-332)."""
        loss, mel = self._step(batch, "valid")
        out = dict(loss=float(loss), mel=mel)
        voc = getattr(self, "vocoder", None)
        if voc is not None and hasattr(voc, "spec2wav") and batch.get("pitches") is not None:
            wavs = []
            for m, f0, n in zip(mel, batch["pitches"], batch["mel_lens"]):
                n = int(n)
                wavs.append(voc.spec2wav(m[:n].T.contiguous(), f0[:n, 0].clone()))
            out["wavs"] = wavs
        return out
