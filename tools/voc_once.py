"""One NSF-HiFiGAN generator pass at the bench shape (B=32, T=4000, config_v1) -- ncu launch-list target."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import Generator, synthetic
dev = torch.device("cuda", 0)
cfg = os.environ.get("CFG", "config_v1.json")
h = json.load(open(os.path.join(ROOT, "tests/golden/nsf_configs", cfg)))
B, T = int(os.environ.get("B", 32)), int(os.environ.get("T", 4000))
gen = Generator(h).to(dev); gen.remove_weight_norm()
gen.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.generator_weights(3, h).items()})
g = torch.Generator().manual_seed(0)
mel = (torch.randn(B, 128, T, generator=g) - 2.5).clamp(-11.5, 2).to(dev)
f0 = (220.0 * 2 ** (0.3 * torch.sin(torch.arange(T) / 50.0))).repeat(B, 1); f0[:, ::5] = 0; f0 = f0.to(dev)
n = int(os.environ.get("N", 2))
gen(mel, f0, seed=1)                      # packs the weights, sizes the work buffers
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    gen(mel, f0, seed=1)
e1.record(); torch.cuda.synchronize()
print(f"done: {e0.elapsed_time(e1) / n:.2f} ms per pass (B={B}, T={T}, {cfg})")
