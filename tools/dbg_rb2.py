import json, os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import Generator
g = dict(np.load("/root/repo/tests/golden/r2_voc_resblock2.npz"))
h = json.loads(str(g["rb2_cfg"]))
sd = {k[len("rb2_sd_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("rb2_sd_")}
dev = torch.device("cuda:0")
mel, f0 = g["rb2_mel"], g["rb2_f0"]
B, T = f0.shape
rng = np.random.RandomState(int(g["rb2_rseed"]))
ri = rng.rand(B, 9).astype(np.float32); nz = rng.randn(B, T * 64, 9).astype(np.float32)
T_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for backend in ("simt", "auto"):
    gen = Generator(h, backend=backend).to(dev); gen.remove_weight_norm(); gen.load_state_dict(sd)
    wav = gen(T_(mel), T_(f0), rand_ini=T_(ri), sine_noise=T_(nz)).cpu().numpy()
    print(backend, rel(wav, g["rb2_wav"]), float(np.abs(wav).max()), np.isnan(wav).any())
# oracle with the same weights
from oracle import nsf_hifigan as ovoc
sdn = {k: v.numpy() for k, v in sd.items()}
try:
    r = ri.copy(); r[:, 0] = 0
    ref = ovoc.generator_forward(sdn, h, mel, f0, r, nz, mode="exact")
    print("oracle vs golden", rel(ref, g["rb2_wav"]))
except Exception as ex:
    print("oracle failed", repr(ex)[:200])
