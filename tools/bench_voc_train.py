"""Vocoder training step on one B200: device time of
  * generator forward + backward on the native nodes (vocoder_train.generator_forward_train), one- and three-product modes,
  * the same generator (UNMODIFIED reference class, oracle/_ref) under torch autograd on the GPU (cuDNN; TF32 as the
    reference's trainer enables it, tools/nsf_hifigan/train.py:28-29, and strict fp32) -- the library path the reference runs,
  * one whole HifiGanTrainer.training_step (discriminator + generator update) with its split,
  * optionally the reference generator forward + backward on the host cores (--cpu).
Synthetic data of configs/vocoder_nsf_hifigan.py's shape: config_v1_256.json, segment 32768 samples (128 frames), batch 20.
Prints one JSON line.  Usage: python tools/bench_voc_train.py [--batch 20] [--frames 128] [--steps 5] [--cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--no-step", action="store_true")
    print(json.dumps(measure(ap.parse_args())))


def measure(a):
    """a: namespace(batch, frames, steps, cpu, no_step) -> dict (see the module docstring)."""
    from fish_diffusion_b200 import Generator, _native as N
    from fish_diffusion_b200 import vocoder_train as VT
    with open(os.path.join(ROOT, "tests", "golden", "nsf_configs", "config_v1_256.json")) as f:
        h = json.load(f)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, T = a.batch, a.frames
    hop = int(np.prod(h["upsample_rates"]))
    S = T * hop
    mel = (torch.randn(B, h["num_mels"], T, device=dev) - 2.5).clamp(-11.5, 2)
    f0 = torch.full((B, T), 220.0, device=dev) * (1 + 0.2 * torch.rand(B, T, device=dev))
    f0[torch.rand(B, T, device=dev) < 0.2] = 0
    gen = Generator(h).to(dev)
    gen_flops_fwd = 614.9e6 * B * T            # 614.9 MFLOP per mel frame (hop 256)
    out = {"config": "config_v1_256.json", "batch": B, "frames": T, "samples": S, "steps": a.steps}
    peak, peak_src = 1400.0, "fallback (B200_PROFILING.md)"
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        with open(pk_path) as f:
            d = json.load(f)
        peak, peak_src = d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json), sustained bf16"
    out["roofline_note"] = {"bound": "tensor", "peak": peak, "unit": "TFLOP/s", "peak_source": peak_src,
                            "algorithmic_flops_per_step": 3 * gen_flops_fwd,
                            "what": "generator forward + backward = 3 x the forward's 614.9 MFLOP per mel frame (SURVEY 8d)"}

    def native_step(cfg):
        for p in gen.parameters():
            p.grad = None
        wav = VT.generator_forward_train(gen, mel, f0, cfg)
        wav.square().mean().backward()

    for prec in ("f16x1", "f16"):
        try:
            cfg = VT.TrainCfg(prec)
            l0 = N.launch_count()
            native_step(cfg)
            launches = N.launch_count() - l0
            ms = timed(lambda: native_step(cfg), a.steps)
            tfl = 3 * gen_flops_fwd / ms / 1e9
            out[f"native_gen_fwd_bwd_{prec}"] = {"ms": ms, "launches_per_step": launches, "algorithmic_tflops": tfl,
                                                 "frac_of_peak_algorithmic": tfl / peak}
        except Exception as e:  # noqa: BLE001
            out[f"native_gen_fwd_bwd_{prec}"] = {"error": repr(e)[:300]}

    try:
        from oracle import ref_loader
        ref = ref_loader.load_reference(with_mel=False)
        rg = ref.nsf.Generator(ref.nsf.AttrDict(h)).to(dev)
        rg.load_state_dict(gen.state_dict())

        def ref_step():
            for p in rg.parameters():
                p.grad = None
            rg(mel, f0).square().mean().backward()

        for tag, tf32 in (("tf32", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            out[f"reference_cudnn_gen_fwd_bwd_{tag}"] = {"ms": timed(ref_step, a.steps)}
        del rg
    except Exception as e:  # noqa: BLE001
        out["reference_cudnn_gen_fwd_bwd"] = {"error": repr(e)[:300]}
    finally:
        torch.backends.cudnn.allow_tf32 = True          # torch's defaults
        torch.backends.cuda.matmul.allow_tf32 = False

    if not a.no_step:
        try:
            from fish_diffusion_b200.vocoder_gan import HifiGanTrainer
            tr = HifiGanTrainer(h, precision="f16x1").to(dev).train()
            t = torch.arange(S, device=dev) / h["sampling_rate"]
            audio = (0.3 * torch.sin(2 * np.pi * 220.0 * t)[None, None] + 0.05 * torch.randn(B, 1, S, device=dev)).contiguous()
            batch = dict(pitches=f0[:, None], audio=audio, audio_lens=torch.full((B,), S, device=dev, dtype=torch.long))
            res = {}
            res["ms"] = timed(lambda: res.__setitem__("loss", tr.training_step(batch)), max(2, a.steps // 2), warmup=1)
            # split: generator forward alone, discriminator update alone
            mels = tr.input_mels(audio, T)
            res["mel_front_end_ms"] = timed(lambda: tr.input_mels(audio, T), a.steps)
            with torch.no_grad():
                y_hat = tr.generate(mels, batch["pitches"]).detach()

            def disc():
                tr.optim_d.zero_grad()
                tr.discriminator_losses(audio, y_hat).backward()
            res["discriminator_fwd_bwd_ms"] = timed(disc, a.steps)
            res["loss"] = {k: float(v) for k, v in res["loss"].items()}
            out["training_step_f16x1"] = res
        except Exception as e:  # noqa: BLE001
            out["training_step_f16x1"] = {"error": repr(e)[:300]}

    if a.cpu:
        try:
            from oracle import ref_loader
            ref = ref_loader.load_reference(with_mel=False)
            threads0 = torch.get_num_threads()
            torch.set_num_threads(os.cpu_count() // 2 or 1)
            rc = ref.nsf.Generator(ref.nsf.AttrDict(h))
            bc = min(B, 2)
            m, f = mel[:bc].cpu(), f0[:bc].cpu()
            rc(m, f).square().mean().backward()
            t0 = time.perf_counter()
            for p in rc.parameters():
                p.grad = None
            rc(m, f).square().mean().backward()
            dt = time.perf_counter() - t0
            out["reference_cpu_gen_fwd_bwd"] = {"ms": dt * 1e3, "batch": bc, "threads": torch.get_num_threads(),
                                                "ms_scaled_to_batch": dt * 1e3 * B / bc}
            torch.set_num_threads(threads0)
        except Exception as e:  # noqa: BLE001
            out["reference_cpu_gen_fwd_bwd"] = {"error": repr(e)[:300]}
    return out


if __name__ == "__main__":
    main()
