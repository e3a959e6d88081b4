"""Two native generator forward + backward passes at the vocoder-training shape (config_v1_256, B=20 x 128 frames) --
the ncu launch-list target of SURVEY.md section 8f N4 (the second pass is the warm one; tools/launch_table.py sums it)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from fish_diffusion_b200 import Generator
from fish_diffusion_b200 import vocoder_train as VT
dev = torch.device("cuda", 0)
h = json.load(open(os.path.join(ROOT, "tests/golden/nsf_configs", "config_v1_256.json")))
B, T = int(os.environ.get("B", 20)), int(os.environ.get("T", 128))
torch.manual_seed(0)
gen = Generator(h).to(dev)
mel = (torch.randn(B, 128, T, device=dev) - 2.5).clamp(-11.5, 2)
f0 = torch.full((B, T), 220.0, device=dev); f0[:, ::5] = 0
cfg = VT.TrainCfg(os.environ.get("PREC", "f16x1"))
for i in range(int(os.environ.get("N", 2))):
    for p in gen.parameters():
        p.grad = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    VT.generator_forward_train(gen, mel, f0, cfg).square().mean().backward()
    e1.record(); torch.cuda.synchronize()
    print(f"pass {i}: {e0.elapsed_time(e1):.2f} ms")
