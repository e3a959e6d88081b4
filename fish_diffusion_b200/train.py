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


class GradSync:
    """Gradient all-reduce that overlaps the backward pass (the role of DDP's per-parameter autograd hooks in the
    reference: configs/_base_/trainers/base.py:30-41, `gradient_as_bucket_view`, 25 MiB buckets).

    The native backward is one autograd node, so hooks cannot see a layer's gradients before the node returns.  Instead
    WaveNetTrainFn.backward hands every finished bucket of `bucket_layers` residual layers (their packed conv /
    conditioner / output-projection weight gradients, 9.4 MB per layer) to `reduce_async`, which starts an NCCL
    all-reduce (average) on the process group's own stream while the remaining layers are still being differentiated;
    `wait` joins them before the gradients are laid out for autograd.  The small remaining gradients (biases, the
    step-embedding MLP, the per-layer diffusion projections, head / tail projections; ~8 M values) are averaged in one
    flat all-reduce after backward (`DenoiserTrainer.step`)."""

    def __init__(self, group=None, bucket_layers=4):
        import torch.distributed as dist
        self.dist, self.group, self.bucket_layers = dist, group, int(bucket_layers)
        self.world = dist.get_world_size(group)
        # NCCL averages inside the collective; other back ends (gloo in the CPU tests) sum and scale afterwards
        self.native_avg = dist.get_backend(group) == "nccl"
        self.handles = []
        self.bytes = 0

    def reduce_async(self, *tensors):
        op = self.dist.ReduceOp.AVG if self.native_avg else self.dist.ReduceOp.SUM
        for t in tensors:
            self.handles.append((self.dist.all_reduce(t, op=op, group=self.group, async_op=True), t))
            self.bytes += t.numel() * t.element_size()

    def wait(self):
        for h, t in self.handles:
            h.wait()
            if not self.native_avg:
                t.div_(self.world)
        self.handles = []


class DenoiserTrainer:
    SYNC_DESCRIPTION = ("bucketed NCCL all-reduce (average) launched from inside the native backward, one bucket per 4 "
                        "residual layers (37.7 MB), overlapping the backward of the layers below; one flat all-reduce "
                        "of the remaining small gradients after backward")

    def __init__(self, diffusion, lr=8e-4, weight_decay=1e-2, betas=(0.9, 0.98), eps=1e-9, clip=0.5,
                 fp16_compress=False, device=None, sync="bucketed", bucket_layers=4):
        """sync: "bucketed" (default, GradSync: all-reduce overlapped with the native backward), "torch" (stock
        DistributedDataParallel: the whole gradient becomes ready at once, so its buckets cannot overlap), "none"
        (no reduction: the compute-only step time of the same process, for the all-reduce split).
        fp16_compress=True reproduces the reference's `ddp_comm_hook=fp16_compress_hook` (trainers/base.py:40) in the
        "torch" mode.  It is off by default: over NVLink 5 the fp32 all-reduce of the 55 M gradients costs ~1.6 ms per
        step while the hook's cast/divide passes cost ~10 ms (measured on 2xB200, profiles/r01_summary.md)."""
        self.diffusion = diffusion
        self.module = TrainStepModule(diffusion)
        self.ddp = None
        self.sync = None
        self.world = 1
        dist_on = (torch.distributed.is_available() and torch.distributed.is_initialized()
                   and torch.distributed.get_world_size() > 1)
        if dist_on and sync != "none":
            self.world = torch.distributed.get_world_size()
            if sync == "torch":
                from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
                from torch.nn.parallel import DistributedDataParallel as DDP
                dev_ids = None if device is None or device.type != "cuda" else [device.index]
                self.ddp = DDP(self.module, device_ids=dev_ids, gradient_as_bucket_view=True, static_graph=True)
                if fp16_compress:
                    self.ddp.register_comm_hook(None, default_hooks.fp16_compress_hook)
            else:
                for p in diffusion.parameters():          # same starting point on every rank (what DDP's constructor does)
                    torch.distributed.broadcast(p.data, src=0)
                self.sync = GradSync(bucket_layers=bucket_layers)
        if hasattr(diffusion, "denoise_fn"):
            diffusion.denoise_fn.grad_sync = self.sync       # None: the backward reduces nothing itself
        params = list(diffusion.parameters())
        # same update rule as the reference's torch.optim.AdamW; the fused (single multi-tensor kernel) implementation
        fused = all(p.is_cuda for p in params)
        self.opt = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps, fused=fused)
        self.clip = clip
        self.last_split_ms = None
        self._ev = None

    def _reduce_rest(self):
        """Average the gradients the native backward did not already reduce, in one flat all-reduce."""
        den = self.diffusion.denoise_fn
        done = set()
        if getattr(den, "_synced_in_backward", False):
            for blk in den.residual_layers:
                done.update(id(p) for p in (blk.conv_layer.conv.weight, blk.conditioner_projection.conv.weight,
                                            blk.output_projection.conv.weight))
        rest = [p.grad for p in self.diffusion.parameters() if p.grad is not None and id(p) not in done]
        if not rest:
            return
        flat = torch._utils._flatten_dense_tensors(rest)
        if self.sync.native_avg:
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.AVG)
        else:
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
            flat.div_(self.world)
        for g, r in zip(rest, torch._utils._unflatten_dense_tensors(flat, rest)):
            g.copy_(r)

    def step(self, features, mel, t=None, noise=None, timing=False):
        """One optimisation step; returns the (local) loss tensor.  timing=True records CUDA events around forward,
        backward (which contains the overlapped all-reduce) and the tail (rest all-reduce, clip, AdamW): the split of the
        previous timed step is in `last_split_ms`."""
        if timing:
            if self._ev is not None:
                e = self._ev
                e[3].synchronize()
                self.last_split_ms = {"forward": e[0].elapsed_time(e[1]), "backward_incl_overlapped_allreduce": e[1].elapsed_time(e[2]),
                                      "rest_allreduce_clip_adamw": e[2].elapsed_time(e[3])}
            self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            self._ev[0].record()
        self.opt.zero_grad(set_to_none=True)
        model = self.ddp if self.ddp is not None else self.module
        loss = model(features, mel, t=t, noise=noise)
        if timing:
            self._ev[1].record()
        loss.backward()
        if timing:
            self._ev[2].record()
        if self.sync is not None:
            self.sync.wait()
            self._reduce_rest()
        if self.clip:
            torch.nn.utils.clip_grad_norm_(self.diffusion.parameters(), self.clip)
        self.opt.step()
        if timing:
            self._ev[3].record()
        return loss.detach()
