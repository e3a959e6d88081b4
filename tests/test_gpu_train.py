"""GPU parity of the training step: native forward + hand-written backward (tap-GEMM dgrad / wgrad kernels) against
the gradients torch autograd produced through the UNMODIFIED reference (tests/golden/train.npz)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import DIFFUSIONS
from gpu_util import dev
from oracle import wavenet as ownet

pytestmark = pytest.mark.gpu


def T_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _build(golden_cfg, name, seed, backend):
    cfg = golden_cfg["WN_" + name.upper()]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", backend=backend, **cfg),
                                 mel_channels=cfg["mel_channels"], noise_loss="smoothed-l1", sampler_interval=10,
                                 spec_min=[-5.0], spec_max=[0.0])).to(dev())
    sd = ownet.make_wavenet_weights(seed, **{k: v for k, v in cfg.items() if k != "dilation_cycle"})
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return diff.train()


@pytest.mark.parametrize("name,seed,backend", [("small", 51, "simt"), ("tc", 52, "simt"), ("tc", 52, "tc")])
def test_train_step_gradients_vs_reference_autograd(golden, golden_cfg, name, seed, backend):
    g = golden("train")
    diff = _build(golden_cfg, name, seed, backend)
    feats = T_(g[f"train_{name}_features"]).requires_grad_(True)
    out = diff.train_step(feats, T_(g[f"train_{name}_mel"]), t=T_(g[f"train_{name}_t"]), noise=T_(g[f"train_{name}_noise"]))
    loss = out["loss"]
    assert loss.requires_grad
    assert abs(float(loss) - float(g[f"train_{name}_loss"])) < 2e-5 * abs(float(g[f"train_{name}_loss"]))
    assert rel_l2(out["epsilon"].detach().cpu().numpy(), g[f"train_{name}_eps"]) < 2e-5
    loss.backward()
    worst = ("", 0.0)
    for k, p in diff.denoise_fn.named_parameters():
        ref = g[f"train_{name}_g_{k}"]
        assert p.grad is not None, k
        e = rel_l2(p.grad.cpu().numpy(), ref)
        if e > worst[1]:
            worst = (k, e)
        assert e < 2e-4, (k, e)
    eg = rel_l2(feats.grad.cpu().numpy(), g[f"train_{name}_gfeatures"])
    print(f"train[{name},{backend}] worst param-grad rel-L2 {worst[1]:.2e} ({worst[0]}), d/dfeatures {eg:.2e}")
    assert eg < 2e-4


def test_optimizer_step_changes_output_and_repacks(golden, golden_cfg):
    """AdamW step on the native module: weights change in place -> packs are rebuilt -> loss decreases on a fixed batch."""
    g = golden("train")
    diff = _build(golden_cfg, "tc", 52, "tc")
    opt = torch.optim.AdamW(diff.parameters(), lr=2e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-2)
    feats, mel = T_(g["train_tc_features"]), T_(g["train_tc_mel"])
    t, noise = T_(g["train_tc_t"]), T_(g["train_tc_noise"])
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = diff.train_step(feats, mel, t=t, noise=noise)["loss"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(diff.parameters(), 0.5)
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


def test_no_grad_train_step_has_no_graph(golden, golden_cfg):
    g = golden("train")
    diff = _build(golden_cfg, "small", 51, "simt")
    with torch.no_grad():
        out = diff.train_step(T_(g["train_small_features"]), T_(g["train_small_mel"]), t=T_(g["train_small_t"]),
                              noise=T_(g["train_small_noise"]))
    assert not out["loss"].requires_grad
    assert abs(float(out["loss"]) - float(g["train_small_loss"])) < 2e-5 * abs(float(g["train_small_loss"]))
