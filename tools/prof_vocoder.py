"""Per-conv-shape device time of the NSF-HiFiGAN generator (CUDA events around every tap-GEMM launch)."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import Generator, synthetic, _native as N
dev = torch.device("cuda", 0)
h = json.load(open(os.path.join(ROOT, "tests/golden/nsf_configs/config_v1.json")))
B, T = int(os.environ.get("B", 32)), int(os.environ.get("T", 4000))
gen = Generator(h).to(dev); gen.remove_weight_norm()
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.generator_weights(3, h).items()})
g = torch.Generator().manual_seed(0)
mel = (torch.randn(B, 128, T, generator=g) - 2.5).clamp(-11.5, 2).to(dev)
f0 = (220.0 * 2 ** (0.3 * torch.sin(torch.arange(T) / 50.0))).repeat(B, 1); f0[:, ::5] = 0; f0 = f0.to(dev)
gen(mel, f0, seed=1); torch.cuda.synchronize()
rec = []
orig = N.conv_cl
def timed(in_planes, w, B_, T_, Cin, Nn, shifts, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(in_planes, w, B_, T_, Cin, Nn, shifts, **kw); e1.record()
    rec.append(((Cin, Nn, len(shifts), T_, "tc" if kw.get("backend", 0) == 0 else "simt"), e0, e1))
N.conv_cl = timed
import fish_diffusion_b200.nsf_hifigan as nh
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gen(mel, f0, seed=1); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
agg = collections.OrderedDict()
for key, a, b in rec:
    d = agg.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b)
conv_ms = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} ms, tap-GEMM convs {conv_ms:.1f} ms, other (source, source convs, conv_post, allocs) {tot-conv_ms:.1f} ms")
for (Cin, Nn, taps, T_, bk), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    flops = 2.0 * B * T_ * Cin * Nn * taps * n
    byt = B * T_ * (4 * Cin + 8 * Nn) * n
    print(f"Cin={Cin:4d} N={Nn:5d} taps={taps:2d} rows/item={T_:8d} {bk:4s} launches={n:3d} {ms:8.2f} ms  {flops/ms/1e9:7.1f} TFLOP/s  ~{byt/ms/1e6:7.0f} GB/s")
