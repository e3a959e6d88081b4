"""GPU parity of the vocoder training path: the CUDA kernels behind
fish_diffusion_b200/vocoder_train.py against torch float64 autograd (node level, tight) and against golden vectors of the
unmodified reference (generator gradients: float64 arbiter; whole training step: HSFHifiGAN.training_step).
Tolerances at generator / step level sit at the reference's own float32 noise, recorded in the goldens
(tests/golden/make_golden_n4.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import n4_util as nu
from fish_diffusion_b200 import _native as N
from gpu_util import dev, planes_to_f64

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _ckpt():
    return torch.load(os.path.join(HERE, "golden", "ref_generator_small.ckpt"), map_location="cpu")["generator"]


def test_lrelu_bwd_kernel():
    torch.manual_seed(0)
    n = 4 * 3 * 1000
    act = torch.randn(n, device=dev())
    act[::7] = 0.0
    grad, addend = torch.randn(n, device=dev()), torch.randn(n, device=dev())
    ap = N.split_nwc(act.view(1, n // 4, 4), N.PREC_F16)
    a = torch.from_numpy(planes_to_f64(ap, N.PREC_F16)).reshape(-1)
    want = torch.where(a > 0, grad.double().cpu(), grad.double().cpu() * 0.1) * 0.5 + addend.double().cpu()
    of = torch.empty(n, device=dev())
    op = torch.empty((2, n), dtype=torch.int16, device=dev())
    N.lrelu_bwd(grad, ap, 0.1, addend=addend, out_f32=of, out_planes=op, scale=0.5, prec=N.PREC_F16)
    assert float((of.double().cpu() - want).abs().max()) < 1e-6
    assert float((torch.from_numpy(planes_to_f64(op, N.PREC_F16)) - want).abs().max()) < 2e-6
    op2 = torch.empty_like(op)
    N.lrelu_bwd(grad, ap, 0.1, out_planes=op2, prec=N.PREC_F16)               # no addend, planes only
    want2 = torch.where(a > 0, grad.double().cpu(), grad.double().cpu() * 0.1)
    assert float((torch.from_numpy(planes_to_f64(op2, N.PREC_F16)) - want2).abs().max()) < 2e-6


def _resblock_reference(x, wb, dil, K, go):
    xr = x.detach().double().cpu().transpose(1, 2).requires_grad_()
    wr = [w.detach().double().cpu().requires_grad_() for w in wb]
    h = xr
    for m, d in enumerate(dil):
        xt = F.conv1d(F.leaky_relu(h, 0.1), wr[4 * m], wr[4 * m + 1], dilation=d, padding=(K * d - d) // 2)
        xt = F.conv1d(F.leaky_relu(xt, 0.1), wr[4 * m + 2], wr[4 * m + 3], padding=(K - 1) // 2)
        h = xt + h
    gr = torch.autograd.grad(h, [xr] + wr, go.double().cpu().transpose(1, 2))
    return h.transpose(1, 2), [gr[0].transpose(1, 2)] + list(gr[1:])


# (C, K, dilations, S): direct weight gradient / 11 taps in two launches / time-folds 2, 4, 8 (the last on the SIMT twin)
NODE_CASES = [(128, 7, (1, 3, 5), 256), (64, 11, (1, 3, 5), 512), (32, 3, (1, 3, 5), 128), (16, 7, (1, 3, 5), 256),
              (8, 3, (1, 3), 64), (256, 3, (1, 3, 5), 64)]


@pytest.mark.parametrize("precision,tol", [("f16", 2e-4), ("f16x1", 4e-2)])
@pytest.mark.parametrize("C,K,dil,S", NODE_CASES)
def test_resblock_node_vs_autograd(C, K, dil, S, precision, tol):
    from fish_diffusion_b200 import vocoder_train as VT
    torch.manual_seed(C * 100 + K)
    cfg = VT.TrainCfg(precision)
    B = 3
    x = torch.randn(B, S, C, device=dev(), requires_grad=True)
    wb = []
    for _ in range(2 * len(dil)):
        wb += [(torch.randn(C, C, K, device=dev()) * (0.9 / (C * K) ** 0.5)).requires_grad_(),
               (torch.randn(C, device=dev()) * 0.1).requires_grad_()]
    go = torch.randn(B, S, C, device=dev()) * 1e-4
    y = VT.ResBlock1Fn.apply(cfg, dil, x, *wb)
    got = torch.autograd.grad(y, [x] + wb, go)
    want_y, want = _resblock_reference(x, wb, dil, K, go)
    errs = [rel(y, want_y)] + [rel(a, b) for a, b in zip(got, want)]
    print(f"ResBlock1Fn[{precision}] C={C} K={K} S={S}: fwd {errs[0]:.2e} dx {errs[1]:.2e} params worst {max(errs[2:]):.2e}")
    assert max(errs) < tol


@pytest.mark.parametrize("Ci,Co,K,S", [(32, 128, 7, 32), (128, 512, 7, 64)])
def test_conv_node_vs_autograd(Ci, Co, K, S):
    from fish_diffusion_b200 import vocoder_train as VT
    torch.manual_seed(Ci)
    cfg = VT.TrainCfg("f16")
    x = torch.randn(2, S, Ci, device=dev(), requires_grad=True)
    w = (torch.randn(Co, Ci, K, device=dev()) * 0.1).requires_grad_()
    b = torch.randn(Co, device=dev()).requires_grad_()
    y = VT.Conv1dFn.apply(cfg, 1, x, w, b)
    go = torch.randn_like(y) * 1e-3
    got = torch.autograd.grad(y, [x, w, b], go)
    xr, wr, br = (t.detach().double().cpu().requires_grad_() for t in (x.transpose(1, 2), w, b))
    h = F.conv1d(xr, wr, br, padding=(K - 1) // 2)
    want = torch.autograd.grad(h, [xr, wr, br], go.double().cpu().transpose(1, 2))
    errs = [rel(y, h.transpose(1, 2)), rel(got[0], want[0].transpose(1, 2)), rel(got[1], want[1]), rel(got[2], want[2])]
    print(f"Conv1dFn {Ci}->{Co}: {['%.2e' % e for e in errs]}")
    assert max(errs) < 1e-4


@pytest.mark.parametrize("Ci,Co,k,u,L", [(128, 64, 8, 4, 16), (64, 32, 8, 4, 64), (32, 16, 4, 2, 128), (16, 8, 4, 2, 256),
                                         (512, 256, 16, 8, 16), (64, 32, 4, 2, 64)])
def test_conv_transpose_node_vs_autograd(Ci, Co, k, u, L):
    from fish_diffusion_b200 import vocoder_train as VT
    torch.manual_seed(Ci + k)
    cfg = VT.TrainCfg("f16")
    p = (k - u) // 2
    x = torch.randn(2, L, Ci, device=dev(), requires_grad=True)
    w = (torch.randn(Ci, Co, k, device=dev()) * 0.1).requires_grad_()
    b = torch.randn(Co, device=dev()).requires_grad_()
    y = VT.ConvTranspose1dFn.apply(cfg, u, p, x, w, b)
    go = torch.randn_like(y) * 1e-3
    got = torch.autograd.grad(y, [x, w, b], go)
    xr, wr, br = (t.detach().double().cpu().requires_grad_() for t in (x.transpose(1, 2), w, b))
    h = F.conv_transpose1d(xr, wr, br, stride=u, padding=p)
    want = torch.autograd.grad(h, [xr, wr, br], go.double().cpu().transpose(1, 2))
    errs = [rel(y, h.transpose(1, 2)), rel(got[0], want[0].transpose(1, 2)), rel(got[1], want[1]), rel(got[2], want[2])]
    print(f"ConvTranspose1dFn {Ci}->{Co} k={k} u={u}: {['%.2e' % e for e in errs]}")
    assert max(errs) < 1e-4


def _draws(seed, shapes):
    rng = np.random.RandomState(seed)
    return [torch.from_numpy((rng.rand(*s) if i == 0 else rng.randn(*s)).astype(np.float32)).to(dev())
            for i, s in enumerate(shapes)]


@pytest.mark.parametrize("name,precision,wav_tol,worst_tol,median_tol", [
    ("n4_gen_init", "f16", 5e-6, 2e-3, 5e-5), ("n4_gen_init", "f16x1", 3e-3, 0.5, 0.1), ("n4_gen", "f16", 3e-4, 5e-2, 1.5e-2)])
def test_generator_gradients_vs_reference_golden(golden, name, precision, wav_tol, worst_tol, median_tol):
    """The CUDA path against the float64 reference Generator (autograd, smooth loss).
      n4_gen_init  the reference's own initialisation: well conditioned (reference float32 vs float64: 7e-8 on the audio,
                   3.6e-6 worst gradient, single LeakyReLU mask flips aside: 5e-4 on one conv) -- the tight pin.  Measured
                   on a B200 (profiles/r02i_n4_gpu_tests.log): three products 3.7e-7 on the audio, worst gradient 6.8e-6,
                   median 1.7e-6; one product 6.9e-5 / 7.2e-2 / 2.0e-2;
      n4_gen       3x those weights: saturated output, chaotic gradient; the reference's own float32 runs sit at worst
                   9.9e-3 / median 2.6e-3 (8 threads) from the arbiter (stored in the golden); measured here on a B200:
                   audio 5.9e-5, worst 1.6e-2, median 5.7e-3.  One-product arithmetic is not compared there."""
    from fish_diffusion_b200 import Generator
    from fish_diffusion_b200 import vocoder_train as VT
    g = golden(name)
    gen = Generator(nu.train_config()).to(dev())
    gen.load_state_dict({k: v * float(g["weight_scale"]) for k, v in _ckpt().items()}, strict=True)
    mel, f0 = torch.from_numpy(g["mels"]).to(dev()), torch.from_numpy(g["pitches"]).to(dev())
    B, S = mel.shape[0], mel.shape[2] * 64
    ri, nz = _draws(nu.SEED_DRAWS + 1, [(B, 9), (B, S, 9)])
    gw = torch.from_numpy(np.random.RandomState(int(g["gw_seed"])).randn(B, 1, S).astype(np.float32)).to(dev())
    launches0 = N.launch_count()
    wav = VT.generator_forward_train(gen, mel, f0, VT.TrainCfg(precision), rand_ini=ri, sine_noise=nz)
    (wav * gw).sum().backward()
    assert N.launch_count() - launches0 > 200          # the native kernels ran (no library fallback)
    e_wav = rel(wav, torch.from_numpy(g["wav"]))
    print(f"[{name},{precision}] wav vs float64 reference {e_wav:.2e} (reference float32: {float(g['noise_f32_wav']):.2e})")
    assert e_wav < wav_tol
    nu.check_gradients(g, [(n, p.grad.cpu().numpy()) for n, p in gen.named_parameters()], "grad_", worst_tol, median_tol,
                       f"[{name},{precision}] generator gradients (CUDA) vs float64 reference")


def test_training_step_vs_reference_golden(golden):
    from fish_diffusion_b200.vocoder_gan import HifiGanTrainer
    g = golden("n4_train")
    tr = HifiGanTrainer(nu.train_config(), precision="f16")
    tr.generator.load_state_dict(_ckpt(), strict=True)
    nu.fill_discriminators(tr.mpd, tr.msd)
    tr.to(dev()).train()
    batch = {k: v.to(dev()) for k, v in nu.make_batch().items()}
    # the native mel front end feeds the generator (train.py:123): against the reference's torchaudio log-mel
    mels = tr.input_mels(batch["audio"], int((batch["audio_lens"] // 64).max()))
    e_mel = float((mels.cpu() - torch.from_numpy(g["mels"])).abs().max())
    print(f"input log-mel: max |native - reference| = {e_mel:.2e}")
    assert mels.shape == g["mels"].shape and e_mel < 2e-3
    B, S = batch["audio"].shape[0], batch["audio"].shape[2]
    # validation_step (train.py:241-270) on the initial weights: inference kernels + native mels against the logged loss
    vri, vnz = _draws(nu.SEED_DRAWS + 2, [(B, 9), (B, S, 9)])
    v = tr.validation_step(batch, rand_ini=vri, sine_noise=vnz)
    print(f"valid_loss {v:.6f} vs reference {float(g['log_valid_loss']):.6f}")
    assert abs(v - float(g["log_valid_loss"])) < 1e-4 * float(g["log_valid_loss"])          # measured 1.6e-6
    ri, nz = _draws(nu.SEED_DRAWS, [(B, 9), (B, S, 9)])
    # the golden is the reference in full float32: keep cuDNN (the discriminators) out of TF32 for this comparison
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out = tr.training_step(batch, rand_ini=ri, sine_noise=nz)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    print(out)
    # measured on a B200: 5.7e-7, 6.6e-6, 1e-7 (profiles/r02i_n4_gpu_tests.log)
    assert abs(out["loss_disc"] - float(g["log_train_loss_disc"])) < 1e-4 * float(g["log_train_loss_disc"])
    assert abs(out["loss_gen"] - float(g["log_train_loss_gen"])) < 1e-4 * float(g["log_train_loss_gen"])
    assert abs(out["envelope"] - float(g["log_train_loss_g_envelope"])) < 1e-4
    for prefix, sub in (("generator", tr.generator), ("mpd", tr.mpd), ("msd", tr.msd)):
        nu.check_gradients(g, [(n, p.grad.cpu().numpy()) for n, p in sub.named_parameters()], f"grad_{prefix}.", 0.5, 0.1,
                           f"training-step gradients of {prefix} (CUDA) vs reference")
    out2 = tr.training_step(batch)                      # a second step with the updated weights and free-running noise
    assert all(np.isfinite(v) for v in out2.values())


def test_real_config_stage_shapes_run():
    """config_v1_256 widths at a short length: every node of the shipped training config runs forward + backward."""
    import json
    from fish_diffusion_b200 import Generator
    from fish_diffusion_b200 import vocoder_train as VT
    with open(os.path.join(HERE, "golden", "nsf_configs", "config_v1_256.json")) as f:
        h = json.load(f)
    torch.manual_seed(0)
    gen = Generator(h).to(dev())
    mel = (torch.randn(2, 128, 8, device=dev()) - 2.5).clamp(-11.5, 2)
    f0 = torch.full((2, 8), 220.0, device=dev())
    wav = VT.generator_forward_train(gen, mel, f0, VT.TrainCfg("f16x1"))
    assert wav.shape == (2, 1, 8 * 256)
    wav.square().mean().backward()
    for n, p in gen.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
