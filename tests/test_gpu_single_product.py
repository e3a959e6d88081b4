"""GPU tests of the single-product GEMM mode ('f16x1' / 'bf16x1': hi planes only, one tensor-core product per k-step).

The mode is the arithmetic of a plain half-precision tensor-core GEMM with fp32 accumulation (BASELINE config #4 asks the
training step in bf16; torch's own CUDA convolutions default to TF32, an 11-bit mantissa like f16x1).  Checked here:
 * the kernels compute EXACTLY sum(hi_a * hi_w) (against float64 on the hi-plane values) on both back ends,
 * the deviation of a whole network evaluation / gradient from the fp32 reference is the expected half-precision level,
   with the thresholds written down.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import DIFFUSIONS, _native as N
from gpu_util import dev, planes_to_f64, tap_gemm_ref
from oracle import wavenet as ownet

pytestmark = pytest.mark.gpu


def hi_to_f64(planes, pc):
    p = planes[0].cpu()
    return p.view(torch.float16 if pc == N.PREC_F16 else torch.bfloat16).to(torch.float64).numpy()


def T_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


@pytest.mark.parametrize("backend", ["simt", "tc"])
@pytest.mark.parametrize("prec", ["f16x1", "bf16x1"])
@pytest.mark.parametrize("case", [(2, 300, 64, 128, [-2, 0, 2]), (1, 500, 16, 16, [-1, 0, 1]),
                                  (2, 257, 128, 256, [-8, 0, 8]), (1, 200, 32, 64, [-1, 0, 1])])
def test_single_product_is_exact_on_hi_planes(case, prec, backend):
    B, T, Ci, Nn, shifts = case
    pc, mma, bk = N.prec_code(prec), N.mma_code(prec), N.backend_code(backend)
    assert mma == pc | N.PREC_SINGLE
    if bk == N.BACKEND_TC and not N.tc_supported_linear(Nn, Ci, len(shifts)):
        pytest.skip("no tensor-core instantiation")
    rng = np.random.RandomState(5 + Ci)
    a = T_(rng.randn(B, T, Ci).astype(np.float32))
    w = T_((rng.randn(Nn, len(shifts) * Ci) / np.sqrt(Ci * len(shifts))).astype(np.float32))
    bias = T_(rng.randn(Nn).astype(np.float32))
    ap, s = N.split_nwc(a, pc), N.pow2_scale(w)
    wp = N.pack_weight(w, pc, s)
    out = torch.empty((B, T, Nn), dtype=torch.float32, device=dev())
    N.conv_cl(ap, wp, B, T, Ci, Nn, shifts, bias=bias, out_f32=out, w_inv_scale=1.0 / s, prec=mma, backend=bk)
    torch.cuda.synchronize()
    b64 = bias.cpu().numpy().astype(np.float64)
    ref_hi = tap_gemm_ref(hi_to_f64(ap, pc), hi_to_f64(wp, pc) / s, shifts, b64)
    ref_full = tap_gemm_ref(planes_to_f64(ap, pc), planes_to_f64(wp, pc) / s, shifts, b64)
    got = out.cpu().numpy()
    assert rel_l2(got, ref_hi) < 2e-6, rel_l2(got, ref_hi)
    # and it really is the reduced arithmetic: half-precision rounding of both operands shows up against the full values
    dev_full = rel_l2(got, ref_full)
    lo, hi = (5e-5, 1e-3) if pc == N.PREC_F16 else (5e-4, 8e-3)
    assert lo < dev_full < hi, dev_full


@pytest.mark.parametrize("name,seed", [("tc", 12), ("full", 0)])
@pytest.mark.parametrize("prec,tol", [("f16x1", 2e-3), ("bf16x1", 2e-2)])
def test_wavenet_forward_single_product_vs_reference(golden, golden_cfg, prec, tol, name, seed):
    g = golden("wavenet")
    cfg = golden_cfg["WN_" + name.upper()]
    sd = ownet.make_wavenet_weights(seed, **{k: v for k, v in cfg.items() if k != "dilation_cycle"})
    from fish_diffusion_b200 import DENOISERS
    outs = {}
    for backend in ("tc", "simt"):
        net = DENOISERS.build(dict(type="WaveNetDenoiser", precision=prec, backend=backend, **cfg)).to(dev())
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        with torch.no_grad():
            outs[backend] = net(T_(g[f"wn_{name}_x"]), torch.tensor([990], device=dev()), T_(g[f"wn_{name}_cond"])).cpu().numpy()
        e = rel_l2(outs[backend], g[f"wn_{name}_y_t990_f64"])
        print(f"wavenet[{name},{prec},{backend}] rel-L2 vs reference fp64 {e:.2e}")
        assert e < tol
    # the two back ends differ only by accumulation order plus re-rounding of slightly different intermediates
    assert rel_l2(outs["tc"], outs["simt"]) < tol


@pytest.mark.parametrize("prec,tol", [("f16x1", 5e-2), ("bf16x1", 1.5e-1)])
def test_train_gradients_single_product_vs_reference_autograd(golden, golden_cfg, prec, tol):
    g = golden("train")
    cfg = golden_cfg["WN_TC"]
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion",
                                 denoiser=dict(type="WaveNetDenoiser", backend="tc", precision=prec, **cfg),
                                 mel_channels=cfg["mel_channels"], noise_loss="smoothed-l1", sampler_interval=10,
                                 spec_min=[-5.0], spec_max=[0.0])).to(dev())
    sd = ownet.make_wavenet_weights(52, **{k: v for k, v in cfg.items() if k != "dilation_cycle"})
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    diff.train()
    out = diff.train_step(T_(g["train_tc_features"]), T_(g["train_tc_mel"]), t=T_(g["train_tc_t"]),
                          noise=T_(g["train_tc_noise"]))
    assert abs(float(out["loss"]) - float(g["train_tc_loss"])) < tol * abs(float(g["train_tc_loss"]))
    out["loss"].backward()
    worst = ("", 0.0)
    for k, p in diff.denoise_fn.named_parameters():
        e = rel_l2(p.grad.cpu().numpy(), g[f"train_tc_g_{k}"])
        if e > worst[1]:
            worst = (k, e)
    print(f"train[tc-config,{prec}] worst param-grad rel-L2 {worst[1]:.2e} ({worst[0]})")
    assert worst[1] < tol, worst
