"""NSF-HiFiGAN training step: the caller of the native generator nodes.

Restates `HSFHifiGAN.training_step` (tools/nsf_hifigan/train.py:114-231) without Lightning (not importable in this
image): one discriminator update and one generator update per batch, manual optimisation, AdamW + ExponentialLR from
the JSON config (train.py:79-96), the SingGAN-style auxiliary losses (L1 STFT at three resolutions, L1 log-mel at three
resolutions, max-pool envelope loss) and the LSGAN / feature-matching terms (models.py:619-649).

Where the arithmetic runs:
  generator forward + backward   native tap-GEMM nodes (vocoder_train.generator_forward_train); the input mel comes from
                                 the native mel front end (mel.get_mel_transform, no gradient needed there)
  discriminators (MPD / MSD, models.py:451-616) and the loss transforms
                                 torch modules with the reference's parameter names (checkpoints of the reference load
                                 with strict=True); their convolutions are grouped / strided / 2-D shapes outside the
                                 tap-GEMM family and stay library calls (cuDNN, cuFFT), like the optimiser.
"""
from __future__ import annotations

import itertools

import torch
from torch import nn
from torch.nn import functional as F
from torch.nn.utils import spectral_norm, weight_norm

from .nsf_hifigan import LRELU_SLOPE, AttrDict, Generator
from .vocoder_train import TrainCfg, generator_forward_train


# ------------------------------------------------------------------------------------------------ discriminators
class DiscriminatorP(nn.Module):
    """One period of the multi-period discriminator (models.py:451-520): the signal folded to [t/period, period] and
    convolved along the first axis only."""

    def __init__(self, period, kernel_size=5, stride=3, use_spectral_norm=False):
        super().__init__()
        self.period = period
        norm_f = spectral_norm if use_spectral_norm else weight_norm
        chans = (1, 32, 128, 512, 1024)
        layers = [norm_f(nn.Conv2d(ci, co, (kernel_size, 1), (stride, 1), padding=(2, 0)))
                  for ci, co in zip(chans[:-1], chans[1:])]
        layers.append(norm_f(nn.Conv2d(1024, 1024, (kernel_size, 1), 1, padding=(2, 0))))
        self.convs = nn.ModuleList(layers)
        self.conv_post = norm_f(nn.Conv2d(1024, 1, (3, 1), 1, padding=(1, 0)))

    def forward(self, x):
        b, c, t = x.shape
        rem = t % self.period
        if rem:
            x = F.pad(x, (0, self.period - rem), "reflect")
            t = x.shape[-1]
        x = x.view(b, c, t // self.period, self.period)
        fmap = []
        for conv in self.convs:
            x = torch.nan_to_num(F.leaky_relu(conv(x), LRELU_SLOPE))
            fmap.append(x)
        x = torch.nan_to_num(self.conv_post(x))
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class DiscriminatorS(nn.Module):
    """One scale of the multi-scale discriminator (models.py:547-577)."""

    SPEC = ((1, 128, 15, 1, 1, 7), (128, 128, 41, 2, 4, 20), (128, 256, 41, 2, 16, 20), (256, 512, 41, 4, 16, 20),
            (512, 1024, 41, 4, 16, 20), (1024, 1024, 41, 1, 16, 20), (1024, 1024, 5, 1, 1, 2))

    def __init__(self, use_spectral_norm=False):
        super().__init__()
        norm_f = spectral_norm if use_spectral_norm else weight_norm
        self.convs = nn.ModuleList([norm_f(nn.Conv1d(ci, co, k, s, groups=g, padding=p)) for ci, co, k, s, g, p in self.SPEC])
        self.conv_post = norm_f(nn.Conv1d(1024, 1, 3, 1, padding=1))

    def forward(self, x):
        fmap = []
        for conv in self.convs:
            x = torch.nan_to_num(F.leaky_relu(conv(x), LRELU_SLOPE))
            fmap.append(x)
        x = torch.nan_to_num(self.conv_post(x))
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class _Multi(nn.Module):
    def _run(self, pairs):
        outs = ([], [], [], [])
        for d, y, y_hat in pairs:
            r, fr = d(y)
            g, fg = d(y_hat)
            for lst, v in zip(outs, (r, g, fr, fg)):
                lst.append(v)
        return outs


class MultiPeriodDiscriminator(_Multi):
    """models.py:523-544."""

    def __init__(self, periods=None):
        super().__init__()
        self.periods = list(periods) if periods is not None else [2, 3, 5, 7, 11]
        self.discriminators = nn.ModuleList([DiscriminatorP(p) for p in self.periods])

    def forward(self, y, y_hat):
        return self._run([(d, y, y_hat) for d in self.discriminators])


class MultiScaleDiscriminator(_Multi):
    """models.py:580-616: the raw signal under spectral norm, then two 2x average-pooled copies under weight norm."""

    def __init__(self):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorS(use_spectral_norm=True), DiscriminatorS(), DiscriminatorS()])
        self.meanpools = nn.ModuleList([nn.AvgPool1d(4, 2, padding=2), nn.AvgPool1d(4, 2, padding=2)])

    def forward(self, y, y_hat):
        pairs = []
        for i, d in enumerate(self.discriminators):
            if i:
                y, y_hat = self.meanpools[i - 1](y), self.meanpools[i - 1](y_hat)
            pairs.append((d, y, y_hat))
        return self._run(pairs)


# ------------------------------------------------------------------------------------------------ losses
def feature_loss(fmap_r, fmap_g):
    """models.py:619-625."""
    tot = 0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            tot = tot + torch.mean(torch.abs(rl - gl))
    return tot * 2


def discriminator_loss(disc_real_outputs, disc_generated_outputs):
    """models.py:628-640 (LSGAN)."""
    tot, r_losses, g_losses = 0, [], []
    for dr, dg in zip(disc_real_outputs, disc_generated_outputs):
        r, g = torch.mean((1 - dr) ** 2), torch.mean(dg ** 2)
        tot = tot + r + g
        r_losses.append(r.item())
        g_losses.append(g.item())
    return tot, r_losses, g_losses


def generator_loss(disc_outputs):
    """models.py:643-649."""
    tot, parts = 0, []
    for dg in disc_outputs:
        l = torch.mean((1 - dg) ** 2)
        parts.append(l)
        tot = tot + l
    return tot, parts


def envelope_loss(y, y_hat, kernel_size=100, stride=50):
    """train.py:98-112: L1 between max-pooled envelopes of the signal and of its negation."""
    env = lambda s: F.max_pool1d(s, kernel_size=kernel_size, stride=stride)
    return F.l1_loss(env(y), env(y_hat)) + F.l1_loss(env(-y), env(-y_hat))


STFT_RESOLUTIONS = ((512, 50, 240), (1024, 120, 600), (2048, 240, 1200))          # train.py:149-153


def stft_loss(y, y_hat):
    """train.py:155-168 (torch.stft with its default rectangular window, as the reference calls it)."""
    tot = 0
    for n_fft, hop, win in STFT_RESOLUTIONS:
        a = torch.view_as_real(torch.stft(y.squeeze(1), n_fft, hop, win, return_complex=True))
        b = torch.view_as_real(torch.stft(y_hat.squeeze(1), n_fft, hop, win, return_complex=True))
        tot = tot + F.l1_loss(a, b)
    return tot / len(STFT_RESOLUTIONS)


def _loss_mel_transform(**kw):
    """Differentiable mel of the loss terms (utils/audio.py:31-60 builds the same torchaudio module)."""
    from torchaudio.transforms import MelSpectrogram
    return MelSpectrogram(center=True, power=1.0, pad_mode="reflect", norm="slaney", mel_scale="slaney", **kw)


def log_mel(transform, audio):
    """train.py:233-240 get_mels."""
    transform = transform.to(audio.device)
    return torch.log(torch.clamp(transform(audio.squeeze(1)), min=1e-5))


def average_gradients(params, group=None):
    """The gradient all-reduce of the reference's DDPStrategy (configs/vocoder_nsf_hifigan.py:25, NCCL over NVLink): one
    flat average of every gradient present -- pass it as `reduce_grads` to HifiGanTrainer.training_step, one process per
    GPU.  (`find_unused_parameters=True` there exists because each backward touches only one of the two networks; here the
    hook is called per network.)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    for g, r in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(r)


# ------------------------------------------------------------------------------------------------ trainer
class HifiGanTrainer(nn.Module):
    """The HSFHifiGAN role (train.py:32-231).  `h`: the JSON config (dict); attribute names follow the reference
    (`generator`, `mpd`, `msd`): the `generator.*`, `mpd.*`, `msd.*` entries of a Lightning checkpoint written by the
    reference load key for key."""

    def __init__(self, h, precision="f16x1", backend="auto", zero_disc_grads=False):
        """zero_disc_grads=False is the reference's behaviour: train.py:114-231 calls `optim_g.zero_grad()` but never
        `optim_d.zero_grad()` under manual optimisation, so the discriminator gradients -- including what the generator
        loss back-propagates into the discriminators -- accumulate over steps.  True clears them before each
        discriminator backward (the conventional HiFi-GAN update)."""
        super().__init__()
        self.zero_disc_grads = bool(zero_disc_grads)
        self.h = h if isinstance(h, AttrDict) else AttrDict(h)
        self.generator = Generator(self.h, precision=precision, backend=backend)
        self.mpd = MultiPeriodDiscriminator(self.h.get("discriminator_periods"))
        self.msd = MultiScaleDiscriminator()
        self.cfg = TrainCfg(precision, backend)
        h = self.h
        from .mel import get_mel_transform
        # generator input: the native mel front end (train.py:43-52 `mel_transform`; no gradient flows through it)
        self.mel_transform = get_mel_transform(sample_rate=h.sampling_rate, n_fft=h.n_fft, hop_length=h.hop_size,
                                               win_length=h.win_size, f_min=h.fmin, f_max=h.fmax, n_mels=h.num_mels)
        # loss side (train.py:55-75): full band on purpose.  A plain list, as in the reference: these transforms are not
        # part of the state_dict and follow the audio's device in log_mel()
        self.multi_scale_mels = [
            _loss_mel_transform(sample_rate=h.sampling_rate, n_fft=n_fft, hop_length=hop, win_length=win, f_min=0,
                                f_max=h.sampling_rate // 2, n_mels=h.num_mels)
            for n_fft, hop, win in ((h.n_fft, h.hop_size, h.win_size), (2048, 270, 1080), (4096, 540, 2160))]
        self.optim_g = self.optim_d = self.sched_g = self.sched_d = None

    def configure_optimizers(self):
        h = self.h
        betas = (h.adam_b1, h.adam_b2)
        # the reference's torch.optim.AdamW (default weight decay); on CUDA its fused single-kernel implementation
        fused = all(p.is_cuda for p in self.parameters())
        self.optim_g = torch.optim.AdamW(self.generator.parameters(), lr=h.learning_rate, betas=betas, fused=fused)
        self.optim_d = torch.optim.AdamW(itertools.chain(self.msd.parameters(), self.mpd.parameters()),
                                         lr=h.learning_rate, betas=betas, fused=fused)
        self.sched_g = torch.optim.lr_scheduler.ExponentialLR(self.optim_g, h.lr_decay)
        self.sched_d = torch.optim.lr_scheduler.ExponentialLR(self.optim_d, h.lr_decay)
        return [self.optim_g, self.optim_d], [self.sched_g, self.sched_d]

    def input_mels(self, y, n_frames):
        """log-mel of the target audio fed to the generator (train.py:123, 233-240)."""
        with torch.no_grad():
            m = self.mel_transform(y.squeeze(1))
            return torch.log(torch.clamp(m, min=1e-5))[:, :, :n_frames]

    def generate(self, mels, pitches, **kw):
        return generator_forward_train(self.generator, mels, pitches, self.cfg, **kw)

    def discriminator_losses(self, y, y_g_hat):
        """train.py:127-135 (the generated signal is detached)."""
        r, g, _, _ = self.mpd(y, y_g_hat.detach())
        loss_f = discriminator_loss(r, g)[0]
        r, g, _, _ = self.msd(y, y_g_hat.detach())
        loss_s = discriminator_loss(r, g)[0]
        return loss_s + loss_f

    def generator_losses(self, y, y_g_hat):
        """train.py:146-216 -> (total, parts)."""
        l_stft = stft_loss(y, y_g_hat)
        l_mel = 0
        for tf in self.multi_scale_mels:
            l_mel = l_mel + F.l1_loss(log_mel(tf, y), log_mel(tf, y_g_hat))
        l_mel = l_mel / len(self.multi_scale_mels)
        l_aux = 0.5 * l_stft + l_mel
        l_env = envelope_loss(y, y_g_hat)
        _, g_f, fr_f, fg_f = self.mpd(y, y_g_hat)
        _, g_s, fr_s, fg_s = self.msd(y, y_g_hat)
        l_fm_f, l_fm_s = feature_loss(fr_f, fg_f), feature_loss(fr_s, fg_s)
        l_gen_f, l_gen_s = generator_loss(g_f)[0], generator_loss(g_s)[0]
        total = l_gen_s + l_gen_f + l_fm_s + l_fm_f + l_env + l_aux * 45
        parts = dict(stft=l_stft, mel=l_mel, envelope=l_env, fm_f=l_fm_f, fm_s=l_fm_s, gen_f=l_gen_f, gen_s=l_gen_s)
        return total, {k: float(v.detach()) for k, v in parts.items()}

    def training_step(self, batch, reduce_grads=None, **gen_kw):
        """batch: dict(pitches [B,1,T] or [B,T], audio [B,1,S], audio_lens [B]) -> dict of losses (train.py:114-231).
        An optional batch["mels"] [B,M,T] (log-mel of the audio, e.g. cached by the data loader) replaces the mel front
        end.  reduce_grads(params): optional gradient all-reduce hook run before each optimiser step (the DDP role of
        configs/vocoder_nsf_hifigan.py:25)."""
        if self.optim_g is None:
            self.configure_optimizers()
        pitches, y = batch["pitches"].float(), batch["audio"].float()
        n_frames = int((batch["audio_lens"] // self.h.hop_size).max())
        mels = batch["mels"][:, :, :n_frames] if batch.get("mels") is not None else self.input_mels(y, n_frames)
        y_g_hat = self.generate(mels, pitches, **gen_kw)
        # ---- discriminators
        loss_d = self.discriminator_losses(y, y_g_hat)
        if self.zero_disc_grads:
            self.optim_d.zero_grad()
        loss_d.backward()
        if reduce_grads is not None:
            reduce_grads(itertools.chain(self.msd.parameters(), self.mpd.parameters()))
        self.optim_d.step()
        # ---- generator
        self.optim_g.zero_grad()
        loss_g, parts = self.generator_losses(y, y_g_hat)
        loss_g.backward()
        if reduce_grads is not None:
            reduce_grads(self.generator.parameters())
        self.optim_g.step()
        return dict(loss_disc=float(loss_d.detach()), loss_gen=float(loss_g.detach()), **parts)

    @torch.no_grad()
    def validation_step(self, batch, **gen_kw):
        """train.py:241-270: masked L1 between the log-mel of the target and of the generated audio -> float.  No gradient
        is needed here, so the generator runs its inference path (fused ResBlock-pair kernels, nsf_hifigan.Generator.forward)
        and both mels come from the native front end."""
        pitches, audios = batch["pitches"].float(), batch["audio"].float()
        mel_lens = batch["audio_lens"] // self.h.hop_size
        n_frames = int(mel_lens.max())
        mels = batch["mels"][:, :, :n_frames] if batch.get("mels") is not None else self.input_mels(audios, n_frames)
        y_g_hat = self.generator(mels, pitches, **gen_kw)
        y_g_hat_mel = self.input_mels(y_g_hat, n_frames)
        mask = (torch.arange(mels.shape[2], device=mels.device)[None, :] < mel_lens[:, None])[:, None].float()
        return float(F.l1_loss(mels * mask, y_g_hat_mel * mask))

    def on_train_epoch_end(self):
        """train.py:225-231: both exponential schedules step once per epoch."""
        self.sched_g.step()
        self.sched_d.step()
