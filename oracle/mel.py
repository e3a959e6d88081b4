"""Oracle: mel / STFT front end (reference fish_diffusion/utils/pitch_adjustable_mel.py:9-96,
fish_diffusion/utils/audio.py:11-18, nsf_hifigan.py:87-107).  TEST INFRASTRUCTURE ONLY.

librosa (pinned 0.9.1 in the reference's pdm.lock) is absent from this image; `slaney_mel_filterbank` restates
librosa.filters.mel(htk=False, norm='slaney') from its published algorithm.  Parity of the filterbank constants
with librosa itself is therefore *unpinned*; it is cross-checked against torchaudio's Slaney filterbank in
tests/test_oracle_golden.py."""
import numpy as np


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) -> float32 [n_mels, 1 + n_fft//2]
    (call site pitch_adjustable_mel.py:46-52)."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def hann_window(n):
    """torch.hann_window(n) (periodic)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def pitch_adjustable_mel(y, sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, f_min=40, f_max=16000,
                         n_mels=128, key_shift=0, speed=1.0, mel_basis=None):
    """PitchAdjustableMelSpectrogram.__call__ (pitch_adjustable_mel.py:33-96).  y [B,N] -> [B,n_mels,frames]."""
    y = np.asarray(y, dtype=np.float64)
    factor = 2 ** (key_shift / 12)
    n_fft_new = int(np.round(n_fft * factor))
    win_new = int(np.round(win_length * factor))
    hop = int(np.round(hop_length * speed))
    if mel_basis is None:
        mel_basis = slaney_mel_filterbank(sample_rate, n_fft, n_mels, f_min, f_max)
    pad = int((win_new - hop) / 2)
    yp = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")
    window = hann_window(win_new)
    if win_new < n_fft_new:  # torch.stft centres a shorter window inside n_fft
        left = (n_fft_new - win_new) // 2
        w = np.zeros(n_fft_new)
        w[left:left + win_new] = window
        window = w
    n_frames = 1 + (yp.shape[1] - n_fft_new) // hop
    idx = np.arange(n_fft_new)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[:, idx] * window[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)               # [B, frames, bins]
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9).transpose(0, 2, 1)   # [B, bins, frames]
    if key_shift != 0:
        size = n_fft // 2 + 1
        resize = mag.shape[1]
        if resize < size:
            mag = np.pad(mag, ((0, 0), (0, size - resize), (0, 0)))
        mag = mag[:, :size, :] * win_length / win_new
    return np.matmul(np.asarray(mel_basis, dtype=np.float64)[None], mag)


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    """audio.py:11-18."""
    return np.log(np.clip(x, clip_val, None) * C)


def wav2spec(y, use_natural_log=True, **kw):
    """NsfHifiGAN.wav2spec (nsf_hifigan.py:87-107) without the resample branch: y [1,N] -> [n_mels, frames]."""
    mel = dynamic_range_compression(pitch_adjustable_mel(y, **kw)[0])
    if not use_natural_log:
        mel = 0.434294 * mel
    return mel
