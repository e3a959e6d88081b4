import sys, torch, importlib.util, os
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge; ge.build()
spec = importlib.util.spec_from_file_location("bt", "/root/repo/tools/bench_train.py"); bt = importlib.util.module_from_spec(spec); spec.loader.exec_module(bt)
from fish_diffusion_b200.train import TrainStepModule
dev = torch.device('cuda:0')
cfg = dict(bt.WN_CFG)
B, T, M, E = 20, 1000, cfg["mel_channels"], cfg["d_encoder"]
g = torch.Generator().manual_seed(100)
feats = torch.randn(2 * B, T, E, generator=g).to(dev); mel = (torch.rand(2 * B, T, M, generator=g) * 5 - 5).to(dev)
t = torch.randint(0, 1000, (2 * B,), generator=g).to(dev); noise = torch.randn(2 * B, M, T, generator=g).to(dev)
def grads(sl_list):
    diff = bt.build(dev, cfg); m = TrainStepModule(diff)
    acc = None
    for sl in sl_list:
        diff.zero_grad(set_to_none=True)
        loss = m(feats[sl], mel[sl], t=t[sl], noise=noise[sl]); loss.backward(); print("loss", sl, float(loss))
        gs = [p.grad.clone() for p in diff.parameters()]
        acc = gs if acc is None else [a + b for a, b in zip(acc, gs)]
    return [a / len(sl_list) for a in acc], [k for k, _ in diff.named_parameters()]
full, names = grads([slice(0, 2 * B)])
halves, _ = grads([slice(0, B), slice(B, 2 * B)])
again, _ = grads([slice(0, 2 * B)])
rows = []
for k, a, b, c in zip(names, full, halves, again):
    e = float((a - b).norm() / a.norm().clamp_min(1e-30)); e2 = float((a - c).norm() / a.norm().clamp_min(1e-30))
    rows.append((e, e2, k))
rows.sort(reverse=True)
for r in rows[:4]: print("%.3e (repeat %.3e) %s" % r)
for r in rows[-6:]: print("%.3e (repeat %.3e) %s" % r)
import collections
byk = collections.defaultdict(list)
for e, e2, k in rows: byk[".".join(k.split(".")[-3:]) if "residual_layers" in k else k].append(e)
for k, v in byk.items(): print("%-60s max %.2e min %.2e" % (k, max(v), min(v)))
