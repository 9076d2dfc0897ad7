"""Griffin-Lim / STFT on the B200 (drop-in for src/data/stft.py::STFT and src/data/audio_processing.py::griffin_lim).

`STFT(640, 160, 640)` keeps the reference constructor / `transform` / `inverse` interface (stft.py:37-129);
`griffin_lim(magnitudes, stft_fn, n_iters)` keeps audio_processing.py:51-68, with the random initial phase drawn on the
device (or injected through `init_angles` for parity tests).  All device work is libvcagan_b200.so (csrc/stft.cu): one
fused frames kernel + one overlap-add kernel per half-iteration, phases never leave the chip."""
import math
from typing import Optional

import torch

from ._lib import lib

N_FFT, HOP = 640, 160


def _check(filter_length, hop_length, win_length, window):
    if (filter_length, hop_length, win_length, window) != (N_FFT, HOP, N_FFT, "hann"):
        raise NotImplementedError("the CUDA STFT is specialised to the reference's configuration "
                                  "(filter_length=640, hop_length=160, win_length=640, window='hann'; vid_aud_grid.py:276)")


def _frames(mode, sig, angles_t, mag_t, spec_out=None):
    B, T, _ = mag_t.shape
    L = HOP * (T - 1)
    frames = torch.empty((B, T, N_FFT), dtype=torch.float32, device=mag_t.device)
    lib().call("vca_gl_frames", mode, sig, angles_t, mag_t, frames, spec_out, B, T, L)
    return frames


def _ola(frames):
    B, T, _ = frames.shape
    L = HOP * (T - 1)
    sig = torch.empty((B, L), dtype=torch.float32, device=frames.device)
    lib().call("vca_gl_ola", frames, sig, B, T, L)
    return sig


class STFT(torch.nn.Module):
    """stft.py:35-133.  transform(x (B,L)) -> (magnitude, phase) each (B,321,frames); inverse(mag, phase) -> (B,1,L')."""

    def __init__(self, filter_length=640, hop_length=160, win_length=640, window='hann'):
        super().__init__()
        _check(filter_length, hop_length, win_length, window)
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window

    def transform(self, input_data):
        x = input_data.reshape(input_data.size(0), -1).contiguous().float()
        B, L = x.shape
        if L % HOP:
            raise ValueError("the CUDA STFT expects num_samples to be a multiple of the hop (160), as produced by inverse()")
        T = L // HOP + 1
        spec = torch.empty((B, T, 321, 2), dtype=torch.float32, device=x.device)
        dummy_mag = torch.zeros((B, T, 321), dtype=torch.float32, device=x.device)
        _frames(1, x, None, dummy_mag, spec)
        re, im = spec[..., 0].transpose(1, 2), spec[..., 1].transpose(1, 2)
        return torch.sqrt(re * re + im * im), torch.atan2(im, re)

    def inverse(self, magnitude, phase):
        mag_t = magnitude.transpose(1, 2).contiguous().float()
        ang_t = phase.transpose(1, 2).contiguous().float()
        return _ola(_frames(0, None, ang_t, mag_t)).unsqueeze(1)

    def forward(self, input_data):
        self.magnitude, self.phase = self.transform(input_data)
        return self.inverse(self.magnitude, self.phase)


# True: one kernel per iteration with the overlap-add inside the CTAs (vca_gl_iter).  Measured SLOWER than the two-kernel path
# (64 clips x 300 frames x 60 iterations: 4.69-4.94 ms vs 3.98-4.10 ms, tools/gl_fused_probe.py): the four conflict-free
# accumulation phases cost barriers and a third of the occupancy, while the 49 MB of frames they save never left L2 anyway.
FUSED_ITERATIONS = False


def griffin_lim(magnitudes, stft_fn=None, n_iters=30, init_angles: Optional[torch.Tensor] = None):
    """audio_processing.py:51-68.  magnitudes (B,321,T') on the GPU -> signal (B, 160*(T'-1)).
    The reference draws the initial phase with (unseeded) numpy on the host; here it is uniform in (-pi, pi] from the
    device Philox stream unless `init_angles` (B,321,T') is given."""
    if not magnitudes.is_cuda:
        raise RuntimeError("griffin_lim needs CUDA tensors: there is no CPU fallback")
    mag_t = magnitudes.transpose(1, 2).contiguous().float()       # frame-major (B,T',321)
    if init_angles is None:
        from . import ops
        u = ops._rng(mag_t.shape, torch.float32, mag_t.device, 2)
        ang_t = u
    else:
        ang_t = init_angles.to(mag_t.device).transpose(1, 2).contiguous().float()
    sig = _ola(_frames(0, None, ang_t, mag_t))
    if n_iters > 0:
        # the magnitudes are read 60 times: reorder their bins once into the lane-major order of the register FFT
        mag_p = torch.empty_like(mag_t)
        lib().call("vca_gl_permute_bins", mag_t, mag_p, mag_t.shape[0] * mag_t.shape[1])
        if not FUSED_ITERATIONS:
            for _ in range(n_iters):
                sig = _ola(_frames(3, sig, None, mag_p))
            return sig
        # one kernel per iteration: the overlap-add happens inside the CTAs, the signal travels as un-normalised sums in two
        # ping-pong buffers (zeroed before each use: neighbouring CTAs add their shared boundary hops atomically)
        B, T = mag_p.shape[0], mag_p.shape[1]
        L = HOP * (T - 1)
        bufs = [torch.empty((B, L), dtype=torch.float32, device=sig.device) for _ in range(2)]
        src, norm = sig, 1
        for i in range(n_iters):
            dst = bufs[i & 1]
            dst.zero_()
            lib().call("vca_gl_iter", src, norm, mag_p, dst, B, T, L)
            src, norm = dst, 0
        sig = torch.empty((B, L), dtype=torch.float32, device=src.device)
        lib().call("vca_gl_normalize", src, sig, B, T, L)
    return sig


# ------------------------------------------------------------------------------------------------------------
# waveform tail / mel front: vid_aud_grid.py:190-240, 270-307; vid_aud_lrs2.py:235-296
# ------------------------------------------------------------------------------------------------------------
LOG1E5 = math.log(1e-5)                 # vid_aud_grid.py:22
DENORM_MUL = -LOG1E5 / 2.0              # denormalize(m) = (m + 1) * (-log1e5 / 2) + log1e5 = m * MUL + ADD
DENORM_ADD = LOG1E5 / 2.0


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float):
    """The (n_mels, 1 + n_fft//2) float32 matrix `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` returns with its
    defaults (Slaney mel scale, triangular filters, area normalisation) -- vid_aud_grid.py:278.  Built once on the
    host at construction time, exactly where the reference builds it."""
    import numpy as np
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz_to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    mels = np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
    mel_f = np.where(mels >= min_log_mel, min_log_hz * np.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)


def _filterbank(x, w, pre, post, pre_mul, pre_add, post_arg):
    """x (B,K,T) fp32, w (K,F) fp32 -> (B,F,T); see vca_filterbank_apply in include/vcagan.h."""
    if not x.is_cuda:
        raise RuntimeError("the filterbank kernels need CUDA tensors: there is no CPU fallback")
    x = x.contiguous().float()
    B, K, T = x.shape
    F = w.shape[1]
    out = torch.empty((B, F, T), dtype=torch.float32, device=x.device)
    lib().call("vca_filterbank_apply", x, w, out, B, K, F, T, pre, post, pre_mul, pre_add, post_arg)
    return out


class TacotronSTFT(torch.nn.Module):
    """vid_aud_grid.py:270-307 (constructed at :38 with 640/160/640, 80 mels, 16 kHz, fmin 55, fmax 7500; LRS 7600)."""

    def __init__(self, filter_length=640, hop_length=160, win_length=640, n_mel_channels=80, sampling_rate=16000,
                 mel_fmin=55.0, mel_fmax=7500.0):
        super().__init__()
        self.n_mel_channels, self.sampling_rate = n_mel_channels, sampling_rate
        self.stft_fn = STFT(filter_length, hop_length, win_length)
        basis = torch.from_numpy(mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax))
        self.register_buffer("mel_basis", basis)                                            # (80, 321), the reference's key
        self.register_buffer("_mel_basis_t", basis.t().contiguous(), persistent=False)      # (321, 80)

    def spectral_normalize(self, magnitudes):            # log(clamp(x, 1e-5)), audio_processing.py:71-78
        return torch.log(torch.clamp(magnitudes, min=1e-5))

    def spectral_de_normalize(self, magnitudes):         # audio_processing.py:81-88
        return torch.exp(magnitudes)

    def mel_spectrogram(self, y):
        """y (B,L) in [-1,1] -> (mel (B,80,frames) log-compressed, magnitudes (B,321,frames)); vid_aud_grid.py:291-307."""
        magnitudes, _ = self.stft_fn.transform(y)
        mel = _filterbank(magnitudes, self._mel_basis_t, 0, 1, 0.0, 0.0, 1e-5)
        return mel, magnitudes

    def mel_to_spec(self, mel):
        """Normalised mel (B,1,80,T) or (B,80,T) -> linear magnitudes (B,321,T) * 1000: the front of inverse_mel
        (vid_aud_grid.py:194-200) as one kernel (denormalize -> exp -> mel_basis matmul -> scale)."""
        m = mel.reshape(-1, self.n_mel_channels, mel.shape[-1])
        return _filterbank(m, self.mel_basis, 1, 0, DENORM_MUL, DENORM_ADD, 1000.0)


def deemphasize(wav: torch.Tensor, coef: float = 0.97, clip: bool = True) -> torch.Tensor:
    """scipy.signal.lfilter([1], [1, -coef], w) per waveform followed by np.clip(-1, 1) (vid_aud_grid.py:205-209,
    230-232), batched on the device: wav (B,L) -> (B,L) fp32."""
    if not wav.is_cuda:
        raise RuntimeError("deemphasize needs CUDA tensors: there is no CPU fallback")
    x = wav.reshape(-1, wav.shape[-1]).contiguous().float()
    y = torch.empty_like(x)
    lo, hi = (-1.0, 1.0) if clip else (-3.0e38, 3.0e38)
    lib().call("vca_deemphasis_clip", x, y, x.shape[0], x.shape[1], float(coef), lo, hi)
    return y.view(wav.shape)


def lrs_denormalize_spec(mag: torch.Tensor) -> torch.Tensor:
    """denormalize -> exp -> * 14 (vid_aud_lrs2.py:261-263, 286-296) as one element-wise kernel."""
    if not mag.is_cuda:
        raise RuntimeError("lrs_denormalize_spec needs CUDA tensors: there is no CPU fallback")
    mag = mag.contiguous().float()
    out = torch.empty_like(mag)
    lib().call("vca_exp_affine", mag, out, mag.numel(), DENORM_MUL, DENORM_ADD, 14.0)
    return out


def inverse_spec(spec, stft: Optional[TacotronSTFT] = None, n_iters: int = 60, lrs: bool = False, init_angles=None):
    """MultiDataset.inverse_spec: GRID vid_aud_grid.py:212-224 (raw magnitudes), LRS vid_aud_lrs2.py:257-272
    (denormalize -> exp -> *14 first).  spec (B,1,321,T) or (1,321,T) -> clipped waveforms (B, 160*(T-1)) on the device."""
    if spec.dim() < 4:
        spec = spec.unsqueeze(0)
    mag = spec.squeeze(1).contiguous().float()
    if lrs:
        mag = lrs_denormalize_spec(mag)
    return deemphasize(griffin_lim(mag, None if stft is None else stft.stft_fn, n_iters, init_angles=init_angles))


def inverse_mel(mel, stft: TacotronSTFT, n_iters: int = 60, init_angles=None):
    """MultiDataset.inverse_mel (vid_aud_grid.py:190-210): normalised mel (B,1,80,T) -> clipped waveforms (B, 160*(T-1))."""
    if mel.dim() < 4:
        mel = mel.unsqueeze(0)
    return deemphasize(griffin_lim(stft.mel_to_spec(mel), stft.stft_fn, n_iters, init_angles=init_angles))
