"""Per-kernel breakdown of one denoiser training step (torch.profiler / CUPTI over the native kernels).

    python tools/prof_train.py [--precision f16|f16x1|bf16x1] [--batch 20] [--frames 1000]
"""
import argparse
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="f16")
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--rows", type=int, default=40)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from fish_diffusion_b200 import _native as N
    from fish_diffusion_b200.train import DenoiserTrainer
    spec = importlib.util.spec_from_file_location("bt", os.path.join(ROOT, "tools", "bench_train.py"))
    bt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bt)
    dev = torch.device("cuda", 0)
    diff = bt.build(dev, dict(bt.WN_CFG, precision=args.precision))
    B, T = args.batch, args.frames
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(B, T, 256, generator=g).to(dev)
    mel = (torch.rand(B, T, 128, generator=g) * 5 - 5).to(dev)
    tr = DenoiserTrainer(diff, device=dev)
    for _ in range(3):
        tr.step(feats, mel)
    torch.cuda.synchronize()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    N.prof_enable(True)
    t0 = time.perf_counter(); e0 = ev()
    tr.opt.zero_grad(set_to_none=True)
    loss = tr.module(feats, mel); e1 = ev(); c1 = time.perf_counter()
    loss.backward(); e2 = ev(); c2 = time.perf_counter()
    torch.nn.utils.clip_grad_norm_(diff.parameters(), 0.5); tr.opt.step(); e3 = ev(); c3 = time.perf_counter()
    torch.cuda.synchronize()
    prof, _ = N.prof_collect()
    N.prof_enable(False)
    print("gpu ms: fwd %.2f bwd %.2f opt %.2f" % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)))
    print("cpu ms: fwd %.2f bwd %.2f opt %.2f" % ((c1 - t0) * 1e3, (c2 - c1) * 1e3, (c3 - c2) * 1e3))
    print("tap-GEMM ms by kind:", {k: (round(v[0], 3), v[1]) for k, v in prof.items()})
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        tr.step(feats, mel)
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 1e3, e.count) for e in p.key_averages() if e.device_time_total > 0 and
            e.device_type == torch.autograd.DeviceType.CUDA]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print("total device kernel time %.2f ms over %d kernel kinds" % (tot, len(rows)))
    for k, ms, n in rows[:args.rows]:
        print("%8.3f ms %5d x  %s" % (ms, n, k[:150]))


if __name__ == "__main__":
    main()
