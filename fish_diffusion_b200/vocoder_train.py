"""Training path of the native NSF-HiFiGAN generator: autograd nodes whose forward AND
backward run on the library's tap-GEMM kernels.

Reference: autograd through fish_diffusion/modules/vocoders/nsf_hifigan/models.py inside
tools/nsf_hifigan/train.py:114-231 (generator step) -- there every conv is a cuDNN forward + dgrad + wgrad and every
LeakyReLU / residual add its own elementwise kernel.  Here:

  ResBlock1Fn      one whole ResBlock1.forward (models.py:103-110, three `x + c2(lrelu(c1(lrelu(x))))` iterations: 89 % of
                   the generator's FLOPs) as ONE autograd node.  Forward = the conv-by-conv tap-GEMM path of
                   nsf_hifigan.Generator with the conv inputs kept as split planes; backward per conv =
                     data gradient    the same tap-GEMM with the transposed, tap-mirrored weights (fd_conv_cl_fwd),
                     weight gradient  both operands read straight from the channels-last planes (fd_wgrad_cl; channel
                                      counts below 64 run on the time-folded view [S/F, F*C] and are un-folded by the
                                      adjoint of nsf_hifigan.fold_conv_weight),
                     bias gradient    column sums (fd_colsum),
                   with the LeakyReLU backward and the residual add fused into one pass (fd_lrelu_bwd).
  Conv1dFn         'same'-padded Conv1d (conv_pre, models.py:362, 417) with the same three kernels.
  ConvTranspose1dFn  the polyphase form of ConvTranspose1d (ups[i], models.py:372-378, 421): forward, data and weight
                   gradients are tap-GEMMs over the [L, u*Co] view of the output.

Gradients inside a node carry a power-of-two scale S (max |S*grad| in [256, 512)) so that they fit the fp16 planes;
everything leaving the node is multiplied by 1/S exactly.  Weight-norm (`weight_g`, `weight_v`) stays under torch
autograd: the nodes take and return the effective weight.  There is no CPU path: the nodes raise on CPU tensors.
"""
from __future__ import annotations

import math

import torch

from . import _native as N

LRELU_SLOPE = 0.1


# ------------------------------------------------------------------------------------------------ packing
def _backend(n_total, k_seg, num_seg, pref="auto"):
    if pref not in ("auto", None):
        return N.backend_code(pref)
    return N.BACKEND_TC if N.tc_supported_linear(n_total, k_seg, num_seg) else N.BACKEND_SIMT


def _pow2_scales(mats, target=64.0):
    """Power-of-two prescales of several weight matrices with ONE device->host transfer (see _native.pow2_scale)."""
    try:
        amax = torch.stack(torch._foreach_norm([m.detach() for m in mats], float("inf"))).tolist()  # one multi-tensor kernel
    except (RuntimeError, TypeError, AttributeError):       # a torch without the inf-norm foreach kernel
        amax = torch.stack([m.detach().abs().max() for m in mats]).tolist()
    out = []
    for m in amax:
        out.append(1.0 if (m == 0.0 or not math.isfinite(m)) else float(2.0 ** math.floor(math.log2(target / m))))
    return out


def conv_offsets(K, d):
    """Row shifts of a 'same'-padded Conv1d with K taps and dilation d (padding (K*d - d)/2, models.py:22-23)."""
    return [(j - (K - 1) // 2) * d for j in range(K)]


def pack_conv_pair(ws, prec):
    """Forward and transposed (data-gradient) packs of several Conv1d weights [Co, Ci, K]:
       fwd  [2][Co][K*Ci]   W[n, j*Ci + c] = w[n, c, j]
       bwd  [2][Ci][K*Co]   Wt[c, j*Co + n] = w[n, c, j]   (applied with the mirrored shifts -off_j)
    -> list of dicts(fwd, bwd, inv)."""
    mats = []
    for w in ws:
        Co, Ci, K = w.shape
        w = w.detach().to(torch.float32)
        mats.append((w.permute(0, 2, 1).reshape(Co, K * Ci).contiguous(), w.permute(1, 2, 0).reshape(Ci, K * Co).contiguous()))
    scales = _pow2_scales([m[0] for m in mats])
    return [dict(fwd=N.pack_weight(f, prec, s), bwd=N.pack_weight(b, prec, s), inv=1.0 / s)
            for (f, b), s in zip(mats, scales)]


# ------------------------------------------------------------------------------------------------ weight gradients
def _wgrad_chunks(rows, cols, shifts, B, T, R, Cc, mma, scale=1.0):
    """scale * sum_{b,t} rows[b,t,r] * cols[b,t+shift_j,c]  for every shift -> fp32 [R, len(shifts), Cc].
    rows [2,B,T,R], cols [2,B,T,Cc] planes with R, Cc multiples of 64; at most 8 shifts per launch."""
    outs = []
    for i in range(0, len(shifts), 8):
        sh = shifts[i:i + 8]
        g = N.wgrad_cl([rows], [cols], [(0, 0, R)], [(0, int(s), 0, Cc) for s in sh], B, T, scale=scale, prec=mma)
        outs.append(g.reshape(R, len(sh), Cc))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


def fold_factor(Co, Ci, T):
    """Time-fold F of a weight-gradient GEMM: the direct kernel needs both channel counts to be multiples of 64; a
    channels-last tensor [T, C] is the same memory as [T/F, F*C], so narrower convs run on the folded view (F = 1 when
    none is needed, 0 if no F <= 8 applies)."""
    for F in (1, 2, 4, 8):
        if (F * Co) % 64 == 0 and (F * Ci) % 64 == 0 and T % F == 0:
            return F
    return 0


def unfold_weight_grad(G, Co, Ci, offs, F, srows):
    """Adjoint of nsf_hifigan.fold_conv_weight: G [F*Co, len(srows), F*Ci] (gradient w.r.t. the folded block-Toeplitz
    weight) -> gradient w.r.t. the tap weights, [Co, len(offs), Ci].  Output sub-step fo of tap j reads folded row
    shift (fo + off_j) // F, input sub-step (fo + off_j) % F."""
    G5 = G.reshape(F, Co, len(srows), F, Ci)
    out = G.new_zeros((Co, len(offs), Ci))
    for j, o in enumerate(offs):
        for fo in range(F):
            out[:, j, :] += G5[fo, :, srows.index((fo + o) // F), (fo + o) % F, :]
    return out


def tap_weight_grad(dy_planes, x_planes, B, T, Co, Ci, offs, mma, scale=1.0):
    """G[n, j, c] = scale * sum_{b,t} dy[b,t,n] * x[b,t+off_j,c]  -> fp32 [Co, len(offs), Ci]: the weight gradient of the
    tap-GEMM y[t] = sum_j W_j x[t + off_j] (rows outside [0,T) read as zero)."""
    F = fold_factor(Co, Ci, T)
    if F == 0:
        raise N.NativeError(f"tap weight gradient: channel counts ({Co}, {Ci}) at length {T} are not supported (both must "
                            "become multiples of 64 under a time-fold of 1, 2, 4 or 8 that divides the length)")
    if F == 1:
        return _wgrad_chunks(dy_planes, x_planes, offs, B, T, Co, Ci, mma, scale)
    srows = sorted({(fo + o) // F for fo in range(F) for o in offs})
    dyf = dy_planes.view(2, B, T // F, F * Co)                                      # same memory, folded rows
    xf = x_planes.view(2, B, T // F, F * Ci)
    G = _wgrad_chunks(dyf, xf, srows, B, T // F, F * Co, F * Ci, mma, scale)        # [F*Co, S, F*Ci]
    return unfold_weight_grad(G, Co, Ci, offs, F, srows)


def conv_weight_grad(dy_planes, x_planes, B, T, Co, Ci, offs, mma, scale=1.0):
    """dW[n, c, j] = scale * sum_{b,t} dy[b,t,n] * x[b,t+off_j,c]  -> fp32 [Co, Ci, K]  (autograd of F.conv1d w.r.t. its
    weight; the permuted view is returned as is -- autograd accumulates it into the [Co, Ci, K] gradient)."""
    return tap_weight_grad(dy_planes, x_planes, B, T, Co, Ci, offs, mma, scale).permute(0, 2, 1)


def _grad_scale(g):
    """Power of two S with max |S*g| in [256, 512): room for the channel sums of the chain below the fp16 maximum while
    the lo planes of gradients 1000x smaller stay out of the subnormals.  One device->host transfer."""
    amax = float(g.detach().abs().max())
    if amax == 0.0 or not math.isfinite(amax):
        return 1.0
    return float(2.0 ** math.floor(math.log2(512.0 / amax)))


class TrainCfg:
    """Arithmetic of the training nodes: storage precision of the planes, GEMM mode ('f16' three products / 'f16x1' one
    product -- the reference trains with TF32 convolutions, tools/nsf_hifigan/train.py:28-29), back end preference."""

    def __init__(self, precision="f16", backend="auto"):
        self.precision = precision
        self.prec = N.prec_code(precision)
        self.mma = N.mma_code(precision)
        self.backend = backend


# ------------------------------------------------------------------------------------------------ ResBlock1
class ResBlock1Fn(torch.autograd.Function):
    """y = ResBlock1(x)  (models.py:103-110) on channels-last tensors.
    apply(cfg, dilations, x_cl [B,S,C], w1_0, b1_0, w2_0, b2_0, w1_1, ...) with effective conv weights [C,C,K]."""

    @staticmethod
    def forward(ctx, cfg, dilations, x_cl, *wb):
        N.require_cuda(x_cl, "x")
        n = len(dilations)
        assert len(wb) == 4 * n
        B, S, C = x_cl.shape
        assert (B * S * C) % 8 == 0 and C % 8 == 0, "ResBlock1Fn: channel count must be a multiple of 8"
        dev = x_cl.device
        prec, mma = cfg.prec, cfg.mma
        i16 = dict(dtype=torch.int16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        x = x_cl.detach().to(torch.float32).contiguous()
        ws = [wb[4 * m + k] for m in range(n) for k in (0, 2)]
        packs = pack_conv_pair(ws, prec)
        K = ws[0].shape[2]
        P0 = N.split_nwc(x, prec)
        PA = torch.empty((2, B, S, C), **i16)                   # planes of lrelu(x_m): input of c1
        N.mrf_finish([P0], PA, in_slope=1.0, scale=1.0, out_slope=LRELU_SLOPE, prec=prec)
        del P0
        saved = []
        cur = x
        for m in range(n):
            p1, p2 = packs[2 * m], packs[2 * m + 1]
            b1 = wb[4 * m + 1].detach().to(torch.float32).contiguous()
            b2 = wb[4 * m + 3].detach().to(torch.float32).contiguous()
            o1, o2 = conv_offsets(K, dilations[m]), conv_offsets(K, 1)
            be = _backend(C, C, K, cfg.backend)
            PB = torch.empty((2, B, S, C), **i16)               # planes of lrelu(c1(.)): input of c2
            N.conv_cl(PA, p1["fwd"], B, S, C, C, o1, bias=b1, w_inv_scale=p1["inv"], out_planes=PB, act=N.ACT_LRELU,
                      act_slope=LRELU_SLOPE, prec=mma, backend=be)
            nxt = torch.empty((B, S, C), **f32)
            PAn = torch.empty((2, B, S, C), **i16) if m < n - 1 else None
            N.conv_cl(PB, p2["fwd"], B, S, C, C, o2, bias=b2, w_inv_scale=p2["inv"], res_f32=cur, out_f32=nxt,
                      out_planes=PAn, act=N.ACT_LRELU, act_slope=LRELU_SLOPE, prec=mma, backend=be)
            saved.append((PA, PB, o1, o2, be))
            cur, PA = nxt, PAn
        ctx.cfg, ctx.packs, ctx.saved, ctx.shape, ctx.K = cfg, packs, saved, (B, S, C), K
        return cur

    @staticmethod
    def backward(ctx, g):
        if ctx.saved is None:
            raise RuntimeError("ResBlock1Fn: backward ran twice (retain_graph is not supported: the saved planes are "
                               "released after the first backward)")
        cfg, packs, saved, K = ctx.cfg, ctx.packs, ctx.saved, ctx.K
        B, S, C = ctx.shape
        prec, mma = cfg.prec, cfg.mma
        dev = g.device
        i16 = dict(dtype=torch.int16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        g = g.detach().to(torch.float32).contiguous()
        Sc = _grad_scale(g)
        inv = 1.0 / Sc
        Gf = g * Sc                                               # fp32 gradient of the residual stream (S-scaled)
        Gp = N.split_nwc(Gf, prec)
        grads = [None] * (4 * len(saved))
        dA = torch.empty((B, S, C), **f32)
        for m in reversed(range(len(saved))):
            PA, PB, o1, o2, be = saved[m]
            p1, p2 = packs[2 * m], packs[2 * m + 1]
            # ---- c2: y = W2 * lrelu(u) + b2, gradient of y is the gradient of the residual stream
            grads[4 * m + 2] = conv_weight_grad(Gp, PB, B, S, C, C, o2, mma, inv)
            grads[4 * m + 3] = N.colsum(Gp, B, S, C, scale=inv, prec=prec)
            N.conv_cl(Gp, p2["bwd"], B, S, C, C, [-o for o in o2], w_inv_scale=p2["inv"], out_f32=dA, prec=mma, backend=be)
            dU = torch.empty((2, B, S, C), **i16)
            N.lrelu_bwd(dA, PB, LRELU_SLOPE, out_planes=dU, prec=prec)
            # ---- c1: u = W1 * lrelu(x) + b1
            grads[4 * m + 0] = conv_weight_grad(dU, PA, B, S, C, C, o1, mma, inv)
            grads[4 * m + 1] = N.colsum(dU, B, S, C, scale=inv, prec=prec)
            N.conv_cl(dU, p1["bwd"], B, S, C, C, [-o for o in o1], w_inv_scale=p1["inv"], out_f32=dA, prec=mma, backend=be)
            del dU
            # ---- d(x_m) = d(x_{m+1}) + lrelu'(x_m) * dA   (fp32 master + planes for the pair below)
            Gn = torch.empty((B, S, C), **f32)
            N.lrelu_bwd(dA, PA, LRELU_SLOPE, addend=Gf, out_f32=Gn, out_planes=Gp if m > 0 else None, prec=prec)
            Gf = Gn
        ctx.saved = None
        dx = Gf * inv if ctx.needs_input_grad[2] else None
        return (None, None, dx) + tuple(grads)


# ------------------------------------------------------------------------------------------------ Conv1d ('same')
class Conv1dFn(torch.autograd.Function):
    """y_cl [B,S,Co] = Conv1d(Ci->Co, K taps, dilation d, 'same')(x_cl [B,S,Ci])  (conv_pre, models.py:362,417).
    apply(cfg, d, x_cl, w [Co,Ci,K], b [Co])."""

    @staticmethod
    def forward(ctx, cfg, d, x_cl, w, b):
        N.require_cuda(x_cl, "x")
        B, S, Ci = x_cl.shape
        Co, _, K = w.shape
        assert Ci % 8 == 0 and Co % 8 == 0, "Conv1dFn: channel counts must be multiples of 8"
        prec, mma = cfg.prec, cfg.mma
        (p,) = pack_conv_pair([w], prec)
        offs = conv_offsets(K, d)
        XP = N.split_nwc(x_cl.detach().to(torch.float32), prec)
        y = torch.empty((B, S, Co), dtype=torch.float32, device=x_cl.device)
        N.conv_cl(XP, p["fwd"], B, S, Ci, Co, offs, bias=b.detach().to(torch.float32).contiguous(), w_inv_scale=p["inv"],
                  out_f32=y, prec=mma, backend=_backend(Co, Ci, K, cfg.backend))
        ctx.cfg, ctx.p, ctx.XP, ctx.offs, ctx.shape = cfg, p, XP, offs, (B, S, Ci, Co, K)
        return y

    @staticmethod
    def backward(ctx, g):
        cfg, p, XP, offs = ctx.cfg, ctx.p, ctx.XP, ctx.offs
        B, S, Ci, Co, K = ctx.shape
        prec, mma = cfg.prec, cfg.mma
        g = g.detach().to(torch.float32).contiguous()
        Sc = _grad_scale(g)
        inv = 1.0 / Sc
        Gp = N.split_nwc(g, prec, scale=Sc)
        gw = conv_weight_grad(Gp, XP, B, S, Co, Ci, offs, mma, inv)
        gb = N.colsum(Gp, B, S, Co, scale=inv, prec=prec)
        dx = None
        if ctx.needs_input_grad[2]:
            dx = torch.empty((B, S, Ci), dtype=torch.float32, device=g.device)
            N.conv_cl(Gp, p["bwd"], B, S, Co, Ci, [-o for o in offs], w_inv_scale=p["inv"] * inv, out_f32=dx, prec=mma,
                      backend=_backend(Ci, Co, K, cfg.backend))
        return None, None, dx, gw, gb


# ------------------------------------------------------------------------------------------------ ConvTranspose1d
def polyphase_taps(k, u, p):
    """Input-row shifts delta of the polyphase form of ConvTranspose1d(k, stride u, padding p): output sample q*u + r
    reads input rows q + delta with kernel tap r + p - delta*u (nsf_hifigan.Generator._pack_convt)."""
    dmin = -((k - 1 - p) // u)
    dmax = (u - 1 + p) // u
    return list(range(dmin, dmax + 1))


def polyphase_weight(w, u, p):
    """ConvTranspose1d weight [Ci, Co, k] -> tap-GEMM matrix W' [u*Co, nd, Ci] (linear in w; its adjoint is
    polyphase_weight_grad)."""
    Ci, Co, k = w.shape
    deltas = polyphase_taps(k, u, p)
    W = w.new_zeros((u, Co, len(deltas), Ci))
    for r in range(u):
        for j, dl in enumerate(deltas):
            kk = r + p - dl * u
            if 0 <= kk < k:
                W[r, :, j, :] = w[:, :, kk].t()
    return W.reshape(u * Co, len(deltas), Ci), deltas


def polyphase_weight_grad(G, Ci, Co, k, u, p):
    """Adjoint of polyphase_weight: G [u*Co, nd, Ci] -> gradient of the ConvTranspose1d weight [Ci, Co, k]."""
    deltas = polyphase_taps(k, u, p)
    G4 = G.reshape(u, Co, len(deltas), Ci)
    out = G.new_zeros((Ci, Co, k))
    for r in range(u):
        for j, dl in enumerate(deltas):
            kk = r + p - dl * u
            if 0 <= kk < k:
                out[:, :, kk] += G4[r, :, j, :].t()
    return out


class ConvTranspose1dFn(torch.autograd.Function):
    """y_cl [B, L*u, Co] = ConvTranspose1d(Ci->Co, k, stride u, padding p)(x_cl [B,L,Ci]) with (k - u) even and
    p = (k - u)/2 (ups[i], models.py:372-378), as a polyphase tap-GEMM.  apply(cfg, u, p, x_cl, w [Ci,Co,k], b [Co])."""

    @staticmethod
    def forward(ctx, cfg, u, p, x_cl, w, b):
        N.require_cuda(x_cl, "x")
        B, L, Ci = x_cl.shape
        _, Co, k = w.shape
        assert k - 2 * p == u, "ConvTranspose1dFn: output length must be L*u (k - 2p == u)"
        assert Ci % 8 == 0 and (u * Co) % 8 == 0
        prec, mma = cfg.prec, cfg.mma
        wd = w.detach().to(torch.float32)
        W3, deltas = polyphase_weight(wd, u, p)                     # [u*Co, nd, Ci]
        nd = len(deltas)
        fwd2d = W3.reshape(u * Co, nd * Ci).contiguous()
        bwd2d = W3.permute(2, 1, 0).reshape(Ci, nd * u * Co).contiguous()      # Wt[c, j*(u*Co) + n]
        (s,) = _pow2_scales([fwd2d])
        pk = dict(fwd=N.pack_weight(fwd2d, prec, s), bwd=N.pack_weight(bwd2d, prec, s), inv=1.0 / s)
        XP = N.split_nwc(x_cl.detach().to(torch.float32), prec)
        y = torch.empty((B, L, u * Co), dtype=torch.float32, device=x_cl.device)
        N.conv_cl(XP, pk["fwd"], B, L, Ci, u * Co, deltas, bias=b.detach().to(torch.float32).repeat(u).contiguous(),
                  w_inv_scale=pk["inv"], out_f32=y, prec=mma, backend=_backend(u * Co, Ci, nd, cfg.backend))
        ctx.cfg, ctx.pk, ctx.XP, ctx.deltas, ctx.shape = cfg, pk, XP, deltas, (B, L, Ci, Co, k, u, p)
        return y.view(B, L * u, Co)

    @staticmethod
    def backward(ctx, g):
        cfg, pk, XP, deltas = ctx.cfg, ctx.pk, ctx.XP, ctx.deltas
        B, L, Ci, Co, k, u, p = ctx.shape
        prec, mma = cfg.prec, cfg.mma
        NN = u * Co
        g = g.detach().to(torch.float32).contiguous().view(B, L, NN)
        Sc = _grad_scale(g)
        inv = 1.0 / Sc
        Gp = N.split_nwc(g, prec, scale=Sc)
        G = tap_weight_grad(Gp, XP, B, L, NN, Ci, deltas, mma, inv)                      # [u*Co, nd, Ci]
        gw = polyphase_weight_grad(G, Ci, Co, k, u, p)
        gb = N.colsum(Gp, B, L, NN, scale=inv, prec=prec).view(u, Co).sum(0)
        dx = None
        if ctx.needs_input_grad[3]:
            dx = torch.empty((B, L, Ci), dtype=torch.float32, device=g.device)
            N.conv_cl(Gp, pk["bwd"], B, L, NN, Ci, [-d for d in deltas], w_inv_scale=pk["inv"] * inv, out_f32=dx,
                      prec=mma, backend=_backend(Ci, NN, len(deltas), cfg.backend))
        return None, None, None, dx, gw, gb


# ------------------------------------------------------------------------------------------------ generator
@torch.no_grad()
def sine_waves(f0_up, sgen, rand_ini=None, noise=None):
    """SineGen.forward (models.py:201-294; its body runs under no_grad in the reference too): f0_up [B,S,1] in Hz ->
    harmonic bank + noise [B,S,H].  Torch ops: in training the H-wide bank is needed as the operand of the l_linear
    weight gradient (the inference kernel fd_sinegen_fwd emits only tanh(l_linear(.))).
    rand_ini [B,H] (column 0 is forced to 0) and noise [B,S,H] may be injected by parity tests."""
    B, S, _ = f0_up.shape
    H = sgen.dim
    # harmonics-major [B,H,S]: the two running sums then scan the contiguous axis (torch's outer-dimension scan of the
    # [B,S,H] layout took 15 ms of a 46 ms generator step at B=20 x 32768 samples)
    f0_t = f0_up.transpose(1, 2)                                     # [B,1,S]
    mult = torch.arange(1, H + 1, device=f0_up.device, dtype=f0_up.dtype).view(1, H, 1)
    rad = ((f0_t * mult) / sgen.sampling_rate) % 1                   # fundamental and overtones
    ri = torch.rand(B, H, device=f0_up.device) if rand_ini is None else rand_ini.to(f0_up).clone()
    ri[:, 0] = 0
    rad[:, :, 0] = rad[:, :, 0] + ri
    over = torch.cumsum(rad, -1) % 1                                 # -1 wherever the running phase wraps
    wrapped = (over[:, :, 1:] - over[:, :, :-1]) < 0
    shift = torch.zeros_like(rad)
    shift[:, :, 1:] = wrapped * -1.0
    sines = torch.sin(torch.cumsum(rad + shift, dim=-1) * 2 * math.pi) * sgen.sine_amp
    uv = (f0_t > sgen.voiced_threshold).to(f0_up.dtype)
    amp = uv * sgen.noise_std + (1 - uv) * sgen.sine_amp / 3
    nz = amp * (torch.randn(B, S, H, device=f0_up.device, dtype=f0_up.dtype) if noise is None else noise.to(sines)).transpose(1, 2)
    return (sines * uv + nz).transpose(1, 2)                         # [B,S,H] view for l_linear


def excitation_conv(har, weight, bias, stride, padding):
    """noise_convs[i] (models.py:380-393, 422): Conv1d(1 -> C, k, stride, padding) of the excitation har [B,S] ->
    channels-last [B, Lo, C], written as windows @ weight^T.  A float32 matmul is exact float32 arithmetic under torch's
    default matmul precision, whereas cuDNN convolutions default to TF32 (measured: 9e-4 on the generated audio)."""
    k = weight.shape[-1]
    hp = torch.nn.functional.pad(har, (padding, padding))
    w = weight[:, 0, :]                                                                   # [C, k]
    n_out = (hp.shape[-1] - k) // stride + 1
    if k % stride == 0 and hp.shape[-1] == (n_out - 1) * stride + k:
        # every shipped config: k = 2*stride (or 1): window t = blocks t .. t + k/stride - 1 of `stride` samples
        blocks = hp.view(hp.shape[0], -1, stride)                                         # [B, Lo + k/s - 1, s]
        out = bias
        for q in range(k // stride):
            out = out + torch.matmul(blocks[:, q:q + n_out], w[:, q * stride:(q + 1) * stride].t())
        return out
    return torch.matmul(hp.unfold(-1, k, stride), w.t()) + bias                           # [B, Lo, k] strided view


def post_conv(x_cl, weight, bias):
    """conv_post (models.py:403, 435): Conv1d(C -> 1, k, 'same') on channels-last x [B,S,C] -> [B,1,S], as one matmul
    x @ W [C,k] followed by the sum of the k shifted columns (exact float32, see excitation_conv)."""
    B, S, C = x_cl.shape
    k = weight.shape[-1]
    half = (k - 1) // 2
    cols = torch.nn.functional.pad(torch.matmul(x_cl, weight[0]), (0, 0, half, half))       # [B, S + 2*half, k]
    out = bias.view(1, 1).expand(B, S)
    for j in range(k):
        out = out + cols[:, j:j + S, j]
    return out[:, None, :]


def generator_forward_train(gen, mel, f0, cfg=None, rand_ini=None, sine_noise=None):
    """Differentiable Generator.forward (models.py:407-438) for the vocoder training step
    (tools/nsf_hifigan/train.py:124, `self.generator(mels, pitches)`): mel [B,M,T], f0 [B,T] or [B,1,T] -> wav [B,1,T*hop].

    conv_pre, every ups[i] and every ResBlock1 (>= 97 % of the FLOPs) run forward AND backward on the native nodes above,
    channels-last throughout; the one-channel ends of the network stay torch ops under autograd: the harmonic source
    (l_linear + tanh over the sine bank), noise_convs[i] (1 -> C strided convs of the excitation, as matmuls), the
    LeakyReLUs between nodes, conv_post (C -> 1, as a matmul) + tanh.  `gen` is a fish_diffusion_b200.Generator with or
    without weight-norm."""
    from torch.nn import functional as Fn
    from .nsf_hifigan import ResBlock1, _effective_weight
    N.require_cuda(mel, "mel")
    if cfg is None:
        cfg = TrainCfg(getattr(gen, "precision", "f16"), "auto")
    if f0.ndim == 2:
        f0 = f0[:, None]
    B, M, T = mel.shape
    hop = 1
    for u in gen.h.upsample_rates:
        hop *= int(u)
    hs = getattr(gen.h, "hop_size", hop)
    if hs != hop:
        raise ValueError(f"hop_size {hs} != product of upsample_rates {hop}")
    f0_up = Fn.interpolate(f0.to(torch.float32), size=T * hop, mode="linear").transpose(1, 2)      # [B,S,1]
    bank = sine_waves(f0_up, gen.m_source.l_sin_gen, rand_ini=rand_ini, noise=sine_noise)
    har = torch.tanh(gen.m_source.l_linear(bank))[:, :, 0]                                        # [B,S]

    x = Conv1dFn.apply(cfg, 1, mel.to(torch.float32).transpose(1, 2).contiguous(), _effective_weight(gen.conv_pre),
                       gen.conv_pre.bias)                                                          # [B,T,C0]
    nk = gen.num_kernels
    for i in range(gen.num_upsamples):
        up, nc = gen.ups[i], gen.noise_convs[i]
        x = Fn.leaky_relu(x, LRELU_SLOPE)
        x = ConvTranspose1dFn.apply(cfg, up.stride[0], up.padding[0], x, _effective_weight(up), up.bias)
        x = x + excitation_conv(har, nc.weight, nc.bias, nc.stride[0], nc.padding[0])
        xs = None
        for j in range(nk):
            rb = gen.resblocks[i * nk + j]
            if not isinstance(rb, ResBlock1):
                raise NotImplementedError("generator_forward_train: only ResBlock1 generators (resblock: \"1\", every "
                                          "shipped training config) have a native backward")
            wb = []
            for c1, c2 in zip(rb.convs1, rb.convs2):
                wb += [_effective_weight(c1), c1.bias, _effective_weight(c2), c2.bias]
            y = ResBlock1Fn.apply(cfg, tuple(rb.dilation), x, *wb)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = Fn.leaky_relu(x)                                              # default slope 0.01 (models.py:434)
    return torch.tanh(post_conv(x, _effective_weight(gen.conv_post), gen.conv_post.bias))
