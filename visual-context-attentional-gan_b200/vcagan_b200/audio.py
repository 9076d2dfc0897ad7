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
    for _ in range(n_iters):
        sig = _ola(_frames(1, sig, None, mag_t))
    return sig
