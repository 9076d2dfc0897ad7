"""Griffin-Lim / STFT CUDA kernels (csrc/stft.cu) against the oracle and the reference's golden vectors
(src/data/stft.py, audio_processing.py).  fp32 FFT vs the reference's dense fp32 DFT: single transforms <= 1e-4
relative (measured ~1e-6); after 8 Griffin-Lim iterations from the same injected initial phase <= 1e-3 (the loop
re-normalises phases every iteration, so rounding differences do not grow)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import vca_oracle as O

pytestmark = pytest.mark.gpu


def test_stft_transform_inverse_match_reference(golden):
    from src.data.stft import STFT
    stft = STFT(640, 160, 640).cuda()
    g = torch.Generator().manual_seed(77)
    sig = torch.randn(2, 160 * 11, generator=g) * 0.1
    mag, ph = stft.transform(sig.cuda())
    assert mag.shape == (2, 321, 12)
    assert rel_l2(mag.cpu(), golden["stft_mag"]) < 1e-4
    # compare phases as phasors where the magnitude is not tiny
    m = torch.from_numpy(golden["stft_mag"]) > 1e-3
    assert float((torch.cos(ph.cpu()) - np.cos(torch.from_numpy(golden["stft_phase"])))[m].abs().max()) < 1e-3
    rec = stft.inverse(torch.from_numpy(golden["stft_mag"]).cuda(), torch.from_numpy(golden["stft_phase"]).cuda())
    assert rec.shape == (2, 1, 160 * 11)
    assert rel_l2(rec.cpu(), golden["stft_rec"]) < 1e-4
    assert float((rec.cpu().squeeze(1) - sig).abs().max()) < 1e-4   # round trip


def test_griffin_lim_matches_reference(golden):
    from src.data.audio_processing import griffin_lim
    mag = torch.from_numpy(golden["gl_mag"]).cuda()
    init = torch.from_numpy(golden["gl_init_phase"]).cuda()
    wav = griffin_lim(mag, None, 8, init_angles=init)
    assert wav.shape == (2, 160 * 11)
    e = rel_l2(wav.cpu(), golden["gl_wav"])
    print("griffin-lim (8 iters) rel err vs reference", e)
    assert e < 1e-3


def test_griffin_lim_full_size_properties():
    """BASELINE config 5 size (64 clips, 321 x 300): output shape, finiteness, and the spectral-convergence property --
    the STFT magnitude of the result must move towards the target as iterations increase."""
    from src.data.audio_processing import griffin_lim
    from src.data.stft import STFT
    g = torch.Generator().manual_seed(5)
    B, T = 64, 300
    # a consistent target: magnitudes of a real signal's STFT
    sig = torch.randn(B, 160 * (T - 1), generator=g).cuda() * 0.1
    stft = STFT(640, 160, 640)
    mag, _ = stft.transform(sig)
    errs = []
    for it in (1, 8, 30):
        wav = griffin_lim(mag, stft, it)
        assert wav.shape == (B, 160 * (T - 1)) and torch.isfinite(wav).all()
        m2, _ = stft.transform(wav)
        errs.append(float((m2 - mag).norm() / mag.norm()))
    print("spectral convergence", errs)
    assert errs[2] < errs[1] < errs[0]
    # oracle cross-check at a size the CPU finishes in seconds
    small = mag[:2, :, :40].contiguous()
    init = (2 * np.pi * torch.rand(2, 321, 40, generator=g) - np.pi).float()
    ref = O.griffin_lim(small.cpu(), init, 5)
    got = griffin_lim(small, stft, 5, init_angles=init)
    assert rel_l2(got.cpu(), ref) < 1e-3


def test_fused_iteration_kernel_matches_two_kernel_path():
    """vca_gl_iter (STFT -> phase x magnitude -> ISTFT -> overlap-add inside the CTAs, un-normalised sums) + vca_gl_normalize
    against gl_frames + gl_ola over a few iterations, incl. a frame count that is not a multiple of the CTA's chunk and clips
    so short that every frame touches the reflected / partially covered ends."""
    from vcagan_b200 import audio as AP
    g = torch.Generator().manual_seed(11)
    for B, T in ((3, 300), (2, 37), (2, 9)):
        mag = (torch.rand(B, 321, T, generator=g) + 0.05).cuda()
        init = ((torch.rand(B, 321, T, generator=g) * 2 - 1) * 3.14159).cuda()
        res = []
        for fused in (False, True):
            AP.FUSED_ITERATIONS = fused
            try:
                res.append(AP.griffin_lim(mag, None, 3, init_angles=init))
            finally:
                AP.FUSED_ITERATIONS = False
        assert torch.isfinite(res[1]).all()
        e = rel_l2(res[1].cpu(), res[0].cpu())
        print("fused vs two-kernel Griffin-Lim", (B, T), e)
        assert e < 1e-4, (B, T, e)
