"""Waveform tail / mel front on the device (csrc/wave_tail.cu) against the reference's
`MultiDataset.inverse_spec / inverse_mel / deemphasize` and `TacotronSTFT.mel_spectrogram` vectors
(tests/golden/golden_tail.npz) and against the oracle at full size.

Tolerances: the de-emphasis recurrence runs in fp64 like scipy's lfilter and is stored as fp32 -> max-abs <= 2e-6 of
full scale; the filterbank matmul is fp32 with a different summation order -> <= 1e-5 relative; whole mel/spec ->
waveform chains go through 60 Griffin-Lim iterations (fp32 FFT vs the reference's dense fp32 DFT) -> <= 1e-4
relative, the bound tests/test_gpu_griffin_lim.py states for the loop itself."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import vca_oracle as O

pytestmark = pytest.mark.gpu


def test_deemphasis_matches_reference(golden_tail):
    from vcagan_b200 import audio
    x = torch.from_numpy(golden_tail["deemph_in"]).cuda()
    y = audio.deemphasize(x)
    assert y.shape == x.shape and y.dtype == torch.float32
    assert np.abs(y.cpu().numpy() - np.clip(golden_tail["deemph_out"], -1, 1)).max() < 2e-6
    raw = audio.deemphasize(x, clip=False)
    assert np.abs(raw.cpu().numpy() - golden_tail["deemph_out"]).max() < 2e-6 * np.abs(golden_tail["deemph_out"]).max()


@pytest.mark.parametrize("B,L", [(1, 1), (3, 4095), (2, 4096), (2, 4097), (64, 47840)])
def test_deemphasis_sizes_and_tile_carries(B, L):
    """Ragged lengths around the 4096-sample tile and the full config-5 batch (64 clips x 47 840 samples)."""
    from vcagan_b200 import audio
    g = torch.Generator().manual_seed(L)
    x = torch.randn(B, L, generator=g) * 0.2
    y = audio.deemphasize(x.cuda()).cpu().numpy()
    assert np.abs(y - O.deemphasize_clip(x.numpy())).max() < 2e-6
    # size-independent property: pre-emphasis (lfilter([1, -0.97], [1]), vid_aud_grid.py:226-228) inverts the filter
    raw = audio.deemphasize(x.cuda(), clip=False).cpu().double()
    back = raw.clone(); back[:, 1:] -= 0.97 * raw[:, :-1]
    assert float((back - x.double()).abs().max()) < 2e-5


def test_mel_basis_and_filterbank(golden_tail):
    from vcagan_b200 import audio
    stft = audio.TacotronSTFT().cuda()
    assert stft.mel_basis.shape == (80, 321)
    assert np.abs(stft.mel_basis.cpu().numpy() - golden_tail["mel_basis"]).max() < 1e-7
    assert "mel_basis" in stft.state_dict()
    spec = stft.mel_to_spec(torch.from_numpy(golden_tail["mel"]).cuda())
    assert spec.shape == (2, 321, 14)
    assert rel_l2(spec.cpu(), golden_tail["mel_to_spec"]) < 1e-5
    mel, mag = stft.mel_spectrogram(torch.from_numpy(golden_tail["melspec_in"]).cuda())
    assert rel_l2(mag.cpu(), golden_tail["melspec_mag"]) < 1e-4
    assert rel_l2(mel.cpu(), golden_tail["melspec_out"]) < 1e-4


def test_inverse_spec_and_mel_match_reference(golden_tail):
    from vcagan_b200 import audio
    gt = golden_tail
    stft = audio.TacotronSTFT().cuda()
    ph = torch.from_numpy(gt["grid_phase"]).cuda()
    wav = audio.inverse_spec(torch.from_numpy(gt["grid_spec"]).cuda(), stft, 60, init_angles=ph)
    assert wav.shape == gt["grid_inverse_spec"].shape
    e1 = rel_l2(wav.cpu(), gt["grid_inverse_spec"])
    wav = audio.inverse_mel(torch.from_numpy(gt["mel"]).cuda(), stft, 60, init_angles=ph)
    e2 = rel_l2(wav.cpu(), gt["grid_inverse_mel"])
    wav = audio.inverse_spec(torch.from_numpy(gt["lrs_spec"]).cuda(), stft, 60, lrs=True, init_angles=ph)
    e3 = rel_l2(wav.cpu(), gt["lrs_inverse_spec"])
    print("inverse_spec / inverse_mel / LRS inverse_spec rel err vs reference:", e1, e2, e3)
    assert max(e1, e2, e3) < 1e-4
    assert float(wav.abs().max()) <= 1.0


def test_filterbank_full_size_against_oracle():
    """Config 5: 64 clips x 300 frames, ragged T (not a multiple of the 32-frame tile)."""
    from vcagan_b200 import audio
    stft = audio.TacotronSTFT(mel_fmax=7600.0).cuda()          # the LRS basis
    basis = O.slaney_mel_basis(fmax=7600.0)
    assert np.abs(stft.mel_basis.cpu().numpy() - basis).max() < 1e-7
    g = torch.Generator().manual_seed(5)
    for B, T in ((64, 300), (3, 37)):
        mel = torch.rand(B, 1, 80, T, generator=g) * 2 - 1
        assert rel_l2(stft.mel_to_spec(mel.cuda()).cpu(), O.mel_to_spec(mel, basis)) < 1e-5
        spec = torch.rand(B, 1, 321, T, generator=g) * 2 - 1
        assert rel_l2(audio.lrs_denormalize_spec(spec.cuda()).cpu(), O.lrs_denormalize_spec(spec)) < 1e-6


def test_synthesize_returns_deemphasised_clipped_wave(state_spec):
    """infer.synthesize = test.py:126-143 including the inverse_spec tail."""
    from conftest import make_state
    from vcagan_b200 import audio, infer, set_precision
    from src.models.visual_front import Visual_front
    from src.models.generator import Decoder, Postnet
    set_precision("fp32")
    mods = {}
    for name, cls in (("v_front", Visual_front), ("gen", Decoder), ("post", Postnet)):
        m = cls().cuda()
        m.load_state_dict(make_state(state_spec, name))
        mods[name] = m
    g = torch.Generator().manual_seed(3)
    vid = torch.randn(1, 1, 20, 112, 112, generator=g).cuda()
    out = infer.synthesize(mods["v_front"], mods["gen"], mods["post"], vid, torch.tensor([20]), n_iters=4)
    assert out["wav"].shape == out["wav_gl"].shape == (1, 160 * 79)
    assert float(out["wav"].abs().max()) <= 1.0
    assert torch.equal(out["wav"], audio.deemphasize(out["wav_gl"]))
