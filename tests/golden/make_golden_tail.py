"""Generate tests/golden/golden_tail.npz: the waveform tail / mel front computed by
the UNMODIFIED reference methods `MultiDataset.inverse_spec / inverse_mel / deemphasize` (src/data/vid_aud_grid.py,
src/data/vid_aud_lrs2.py) and `TacotronSTFT.mel_spectrogram`.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_tail.py
Shims: librosa and matplotlib are not installed.  librosa.util gets the 3 no-op functions of SURVEY.md section 8(c);
`librosa.filters.mel` -- the one librosa function whose arithmetic matters here -- is replaced by
oracle.vca_oracle.slaney_mel_basis, so the mel BASIS itself stays unpinned (stated in the oracle header) while
everything the reference does WITH it (denormalise, exp, matmul, scaling, Griffin-Lim, lfilter, clip) is the reference's
own code.  The dataset objects are created without running __init__ (which walks the dataset directory); only the
attributes the methods read are set.
"""
import os, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("VCA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from oracle import vca_oracle as O  # noqa: E402

lib = types.ModuleType("librosa"); util = types.ModuleType("librosa.util"); filt = types.ModuleType("librosa.filters")
util.pad_center = lambda data, size, **k: data
util.tiny = lambda x: np.finfo(np.float32).tiny
util.normalize = lambda x, norm=None, **k: x
filt.mel = lambda sr, n_fft, n_mels, fmin, fmax: O.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
lib.util = util; lib.filters = filt
mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt; mpl.use = lambda *a, **k: None
sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt, "matplotlib": mpl,
                    "matplotlib.pyplot": plt})

import src.data.vid_aud_grid as grid  # noqa: E402
import src.data.vid_aud_lrs2 as lrs2  # noqa: E402


def fixed_phase(shape, gen):
    """Patch numpy's RNG the way make_golden.py does so griffin_lim's initial phase (audio_processing.py:59) is known."""
    init = (2 * np.pi * torch.rand(*shape, generator=gen) - np.pi).float()
    phase01 = ((init.numpy().astype(np.float64)) / (2 * np.pi)) % 1.0
    used = np.angle(np.exp(2j * np.pi * phase01)).astype(np.float32)
    return phase01, used


def main():
    g = torch.Generator().manual_seed(4321)
    B, T = 2, 14
    ds = object.__new__(grid.MultiDataset)
    ds.f_min, ds.f_max = 55., 7500.
    stft = grid.TacotronSTFT(filter_length=640, hop_length=160, win_length=640, n_mel_channels=80, sampling_rate=16000,
                             mel_fmin=55., mel_fmax=7500.)
    dl = object.__new__(lrs2.MultiDataset)
    out = {}

    wav_in = (torch.randn(3, 5000, generator=g) * 0.3).numpy()
    out["deemph_in"] = wav_in
    out["deemph_out"] = np.stack([ds.deemphasize(w) for w in wav_in])            # float64, not clipped

    spec = torch.rand(B, 1, 321, T, generator=g) * 0.02                           # GRID: raw magnitudes
    phase01, used = fixed_phase((B, 321, T), g)
    orig = np.random.rand
    np.random.rand = lambda *a: phase01
    out["grid_spec"], out["grid_phase"] = spec, used
    out["grid_inverse_spec"] = ds.inverse_spec(spec, stft)

    mel = torch.rand(B, 1, 80, T, generator=g) * 2 - 1
    out["mel"] = mel
    out["grid_inverse_mel"] = ds.inverse_mel(mel, stft)
    m = stft.spectral_de_normalize(ds.denormalize(mel)).transpose(2, 3).contiguous()
    out["mel_to_spec"] = (torch.matmul(m, stft.mel_basis).transpose(2, 3).squeeze(1) * 1000)

    lspec = torch.rand(B, 1, 321, T, generator=g) * 2 - 1                         # LRS: normalised log spectrogram
    out["lrs_spec"] = lspec
    out["lrs_inverse_spec"] = dl.inverse_spec(lspec, stft)
    np.random.rand = orig

    y = torch.clamp(torch.randn(2, 160 * 9, generator=g) * 0.2, -1, 1)
    melspec, mags = stft.mel_spectrogram(y)
    out.update(melspec_in=y, melspec_out=melspec, melspec_mag=mags, mel_basis=stft.mel_basis)

    path = os.path.join(HERE, "golden_tail.npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in out.items()})
    print("wrote", len(out), "arrays;", os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
