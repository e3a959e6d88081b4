#!/usr/bin/env python
"""Denoiser training-step benchmark (BASELINE config #4: WaveNet train_step, fwd + bwd + NCCL all-reduce under DDP).

  python tools/bench_train.py [--batch 20] [--frames 1000] [--steps 10] [--warmup 3]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

Also the 2-rank gradient-equality check (--check-ddp): gradients of N ranks x per-rank batch b, averaged by DDP,
equal the gradients of one process on the concatenated global batch (fp16-compress hook off for the check).
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WN_CFG = dict(mel_channels=128, d_encoder=256, residual_channels=512, residual_layers=20, use_linear_bias=True,
              dilation_cycle=4)


def build(dev, cfg, seed=0):
    from fish_diffusion_b200 import DIFFUSIONS, synthetic
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **cfg),
                                 mel_channels=cfg["mel_channels"], noise_loss="smoothed-l1", sampler_interval=10,
                                 spec_min=[-5.0], spec_max=[0.0])).to(dev)
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.wavenet_weights(seed, **cfg).items()})
    return diff.train()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check-ddp", action="store_true")
    ap.add_argument("--small", action="store_true", help="reduced network (tests)")
    ap.add_argument("--precision", default="f16", help="f16 | bf16 (three split products) or f16x1 | bf16x1 (one product)")
    ap.add_argument("--fp16-hook", action="store_true", help="use the reference's fp16_compress_hook (sync=torch)")
    ap.add_argument("--sync", default="bucketed", choices=["bucketed", "torch", "none"],
                    help="gradient reduction: overlapped buckets from inside the native backward / stock DDP / none")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from fish_diffusion_b200 import _native as N
    from fish_diffusion_b200.dist import init_process_group, max_over_ranks
    from fish_diffusion_b200.train import DenoiserTrainer

    rank, world, local = init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = dict(WN_CFG, residual_channels=128, d_encoder=64, mel_channels=64, residual_layers=3) if args.small else WN_CFG
    cfg = dict(cfg, precision=args.precision)
    B, T, M, E = args.batch, args.frames, cfg["mel_channels"], cfg["d_encoder"]
    g = torch.Generator().manual_seed(100)
    feats_all = torch.randn(world * B, T, E, generator=g)
    mel_all = torch.rand(world * B, T, M, generator=g) * 5 - 5
    t_all = torch.randint(0, 1000, (world * B,), generator=g)
    noise_all = torch.randn(world * B, M, T, generator=g)
    sl = slice(rank * B, (rank + 1) * B)
    feats, mel, t, noise = (a[sl].to(dev) for a in (feats_all, mel_all, t_all, noise_all))

    if args.check_ddp:
        diff = build(dev, cfg)
        tr = DenoiserTrainer(diff, fp16_compress=False, device=dev, sync=args.sync, clip=0)
        model = tr.ddp if tr.ddp is not None else tr.module
        model(feats, mel, t=t, noise=noise).backward()
        if tr.sync is not None:
            tr.sync.wait()
            tr._reduce_rest()
        ref = build(dev, cfg)
        DenoiserTrainer(ref).module(feats_all.to(dev), mel_all.to(dev), t=t_all.to(dev), noise=noise_all.to(dev)).backward()
        worst = 0.0
        for (k, p), (_, q) in zip(diff.named_parameters(), ref.named_parameters()):
            e = float((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-30))
            worst = max(worst, e)
        worst = max_over_ranks(worst, dev)
        if rank == 0:
            print(json.dumps({"check": "ddp_gradient_equality", "world": world, "per_rank_batch": B, "frames": T,
                              "worst_rel_l2": worst, "ok": worst < 1e-3}))
        if world > 1:
            torch.distributed.destroy_process_group()
        sys.exit(0 if worst < 1e-3 else 1)

    diff = build(dev, cfg)
    tr = DenoiserTrainer(diff, device=dev, fp16_compress=args.fp16_hook, sync=args.sync)
    for _ in range(args.warmup):
        tr.step(feats, mel)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(feats, mel)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    # forward-only time for the split
    with torch.no_grad():
        diff.train_step(feats, mel)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            diff.train_step(feats, mel)
        e1.record()
        torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / args.steps
    flops = 3 * 95.159e6 * B * T          # fwd + 2x bwd, algorithmic (95.159 MFLOP / position forward)
    if rank == 0:
        print(json.dumps({
            "metric": "denoiser_train_step", "n_gpus": world, "per_gpu_batch": B, "frames": T, "ms_per_step": ms,
            "samples_per_sec": world * B / (ms * 1e-3), "mel_frames_per_sec": world * B * T / (ms * 1e-3),
            "fwd_only_ms": fwd_ms, "algorithmic_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
            "gpu_launches_per_step": (N.launch_count() - l0) // (2 * args.steps) if False else None,
            "loss": float(loss), "optimizer": "AdamW(8e-4, wd 1e-2, betas (0.9,0.98), eps 1e-9), clip 0.5",
            "ddp": (args.sync + (", fp16_compress_hook" if args.fp16_hook else "")) if world > 1 else "single process",
            "dtype": ("%s operands, fp32 accumulate (single tcgen05 product)" % args.precision[:-2]) if args.precision.endswith("x1") else "f32 (3x %s split-product tcgen05)" % args.precision, "precision": args.precision, "data": "synthetic"}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
