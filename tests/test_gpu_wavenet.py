"""GPU parity: native WaveNet + samplers through the reference-facing classes against the golden vectors of the
unmodified reference and the float64 oracle (tolerance of north_star: 1e-3 relative fp32; we assert far tighter)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import DIFFUSIONS, WaveNet
from gpu_util import dev
from oracle import sampler as osamp
from oracle import wavenet as ownet

pytestmark = pytest.mark.gpu


def wn_weights(seed, cfg):
    return ownet.make_wavenet_weights(seed, **{k: v for k, v in cfg.items() if k != "dilation_cycle"})


def build_net(cfg, sd, **kw):
    net = WaveNet(**cfg, **kw).to(dev())
    r = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    return net.eval()


def T_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


@pytest.mark.parametrize("name,seed,backends", [("small", 11, ["simt"]), ("tc", 12, ["simt", "tc"]),
                                                ("nobias", 13, ["simt"]), ("full", 0, ["tc", "simt"])])
@pytest.mark.parametrize("prec", ["f16", "bf16"])
def test_wavenet_forward_vs_reference(golden, golden_cfg, name, seed, backends, prec):
    g = golden("wavenet")
    cfg = golden_cfg["WN_" + name.upper()]
    sd = wn_weights(seed, cfg)
    x, cond = g[f"wn_{name}_x"], g[f"wn_{name}_cond"]
    tol = 2e-5 if prec == "f16" else 3e-4
    for backend in backends:
        net = build_net(cfg, sd, precision=prec, backend=backend)
        with torch.no_grad():
            y = net(T_(x), torch.tensor([990], device=dev()), T_(cond)).cpu().numpy()
        e32, e64 = rel_l2(y, g[f"wn_{name}_y_t990"]), rel_l2(y, g[f"wn_{name}_y_t990_f64"])
        print(f"wavenet[{name},{prec},{backend}] rel-L2 vs ref fp32 {e32:.2e}, vs ref fp64 {e64:.2e}")
        assert e64 < tol and e32 < tol
        if name in ("small", "tc"):
            B = x.shape[0]
            with torch.no_grad():
                ys = net(T_(x), T_(np.array([17.0, 503.25][:B], dtype=np.float32)), T_(cond)).cpu().numpy()
                m = g[f"wn_{name}_masks"]
                ym = net(T_(x), torch.tensor([40], device=dev()), T_(cond), x_masks=T_(m), cond_masks=T_(m)).cpu().numpy()
                y4 = net(T_(x)[:, None], torch.tensor([990], device=dev()), T_(cond))
            assert rel_l2(ys, g[f"wn_{name}_y_stepsB"]) < tol
            assert rel_l2(ym, g[f"wn_{name}_y_masked_t40"]) < tol
            assert np.all(ym[0, :, -7:] == 0) and np.all(ym[1, :, -19:] == 0)
            assert y4.shape == (B, 1, cfg["mel_channels"], x.shape[2])
            assert np.array_equal(y4[:, 0].cpu().numpy(), y)


def test_wavenet_zero_init_output_projection_gives_bias_only():
    """A freshly constructed WaveNet outputs only its bias (wavenet.py:192) -- same here."""
    net = WaveNet(mel_channels=64, d_encoder=64, residual_channels=128, residual_layers=2, dilation_cycle=2).to(dev())
    x = torch.randn(1, 64, 33, device=dev())
    with torch.no_grad():
        y = net(x, torch.tensor([5], device=dev()), torch.randn(1, 64, 33, device=dev()))
    want = net.output_projection.conv.bias.detach()[None, :, None].expand_as(y)
    assert torch.allclose(y, want, atol=1e-6)


def test_wavenet_repacks_when_weights_change(golden_cfg):
    cfg = golden_cfg["WN_TC"]
    net = build_net(cfg, wn_weights(1, cfg))
    x, c = torch.randn(1, 64, 50, device=dev()), torch.randn(1, 64, 50, device=dev())
    t = torch.tensor([100], device=dev())
    with torch.no_grad():
        y1 = net(x, t, c).clone()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in wn_weights(2, cfg).items()})
        y2 = net(x, t, c).clone()
    ref2 = ownet.wavenet_forward(wn_weights(2, cfg), x.cpu().numpy(), np.array([100]), c.cpu().numpy(), dilation_cycle=4)
    assert rel_l2(y2.cpu().numpy(), ref2) < 2e-5 and rel_l2(y1.cpu().numpy(), ref2) > 1e-2


def test_reference_layout_forward_is_differentiable(golden_cfg):
    """WaveNet.forward([B,M,T]) with grad enabled goes through the native backward (never silently detached): gradients
    reach the parameters, the conditioner AND x; masks act as in the reference forward (wavenet.py:217-221,233-234)."""
    cfg = golden_cfg["WN_TC"]
    net = build_net(cfg, wn_weights(1, cfg))
    x = torch.randn(1, 64, 20, device=dev())
    c = torch.randn(1, 64, 20, device=dev(), requires_grad=True)
    y = net(x, torch.tensor([1], device=dev()), c)
    assert y.requires_grad and y.shape == (1, 64, 20)
    y.square().mean().backward()
    assert c.grad is not None and torch.isfinite(c.grad).all() and float(c.grad.abs().max()) > 0
    assert net.input_projection.conv.weight.grad is not None
    with torch.no_grad():
        y0 = net(x, torch.tensor([1], device=dev()), c)
    assert torch.allclose(y0, y.detach(), rtol=1e-5, atol=1e-6)      # training and inference forwards agree
    # masks under grad: same values as the masked inference forward, zero gradient into masked conditioner rows
    m = torch.zeros(1, 20, dtype=torch.bool, device=dev())
    m[0, 13:] = True
    xg = x.clone().requires_grad_(True)
    c2 = c.detach().clone().requires_grad_(True)
    ym = net(xg, torch.tensor([1], device=dev()), c2, x_masks=m, cond_masks=m)
    with torch.no_grad():
        ym0 = net(x, torch.tensor([1], device=dev()), c, x_masks=m, cond_masks=m)
    assert torch.allclose(ym0, ym.detach(), rtol=1e-5, atol=1e-6)
    assert float(ym.detach()[0, :, 13:].abs().max()) == 0.0
    ym.square().mean().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and float(xg.grad.abs().max()) > 0
    assert float(c2.grad[0, :, 13:].abs().max()) == 0.0 and float(c2.grad[0, :, :13].abs().max()) > 0


# ------------------------------------------------------------------ samplers, noise injected
def _build_diffusion(golden_cfg, pred, interval, backend="simt"):
    cfg = golden_cfg["WN_SMALL"]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", backend=backend, **cfg),
                                 mel_channels=16, noise_schedule="linear", timesteps=1000, max_beta=0.01,
                                 noise_loss="smoothed-l1", sampler_interval=interval, spec_min=[-5.0], spec_max=[0.0],
                                 noise_predictor=pred)).to(dev())
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in wn_weights(21, cfg).items()})
    return diff.eval()


@pytest.mark.parametrize("pred,interval,skip", [("naive", 100, 0), ("naive", 50, 900), ("plms", 100, 0),
                                                ("plms", 50, 900), ("unipc", 100, 0)])
def test_sampler_vs_reference_with_injected_noise(golden, golden_cfg, pred, interval, skip):
    g = golden("sampler")
    key = f"samp_{pred}_i{interval}_s{skip}"
    noises = [T_(g[key + f"_noise{j}"]) for j in range(int(g[key + "_nnoise"]))]
    diff = _build_diffusion(golden_cfg, pred, interval)
    kw = {}
    if skip:
        kw["original_mel"] = T_(np.transpose(g[key + "_original_mel"], (0, 2, 1)))
    mel = diff(T_(g["samp_features"]), sampler_interval=interval, skip_steps=skip, noise_predictor=pred,
               x_T=noises[0], step_noises=noises[1:], **kw)
    e = rel_l2(mel.cpu().numpy(), g[key + "_mel"])
    print(f"sampler[{pred},i{interval},s{skip}] rel-L2 vs reference {e:.2e}")
    assert mel.shape == g[key + "_mel"].shape
    assert e < 1e-4


def test_sampler_free_running_philox_statistics(golden, golden_cfg):
    """No injection: in-kernel Philox.  Output must be finite, inside the denormalised clip range, and differ
    between two calls (the offset advances) while being reproducible under the same torch seed."""
    diff = _build_diffusion(golden_cfg, "naive", 100)
    feats = T_(golden("sampler")["samp_features"])
    torch.manual_seed(5)
    a = diff(feats, sampler_interval=100, noise_predictor="naive")
    b = diff(feats, sampler_interval=100, noise_predictor="naive")
    assert torch.isfinite(a).all() and float(a.min()) >= -5.0 - 1e-4 and float(a.max()) <= 1e-4
    assert not torch.equal(a, b)
    torch.manual_seed(5)          # the Philox key is drawn from torch's generator: re-seeding reproduces both calls
    a2 = diff(feats, sampler_interval=100, noise_predictor="naive")
    b2 = diff(feats, sampler_interval=100, noise_predictor="naive")
    assert torch.equal(a, a2) and torch.equal(b, b2)


def test_randn_kernel_moments():
    from fish_diffusion_b200 import _native as N
    out = torch.empty(1 << 22, device=dev())
    N.check(N.lib().fd_randn(N.ptr(out), out.numel(), 1234, 0, 0, N.stream_ptr(dev())), "randn")
    m, s = float(out.mean()), float(out.std())
    k = float(((out - m) ** 4).mean() / s ** 4)
    assert abs(m) < 3e-3 and abs(s - 1) < 3e-3 and abs(k - 3) < 0.05


def test_train_step_forward_vs_reference(golden, golden_cfg):
    g = golden("sampler")
    diff = _build_diffusion(golden_cfg, "naive", 10)
    out = diff.train_step(T_(g["samp_features"]), T_(g["train_mel"]), t=T_(g["train_t"]), noise=T_(g["train_noise"]))
    assert out["loss"].requires_grad                      # grad mode: the loss carries the native backward
    assert rel_l2(out["noised_mels"].detach().cpu().numpy(), g["train_noised"]) < 1e-6
    assert rel_l2(out["epsilon"].detach().cpu().numpy(), g["train_eps"]) < 2e-5
    assert abs(float(out["loss"]) - float(g["train_loss"])) < 2e-5 * abs(float(g["train_loss"]))
