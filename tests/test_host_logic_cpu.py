"""Host-side logic of the product path that needs no GPU: the tables it computes on the host at construction time,
checked against the oracle's independent restatement (and the live third-party library where installed)."""
import numpy as np
import pytest

from oracle import vca_oracle as O


@pytest.mark.parametrize("n_in,n_out", [(136, 112), (80, 112), (112, 112), (300, 112), (57, 112), (100, 112), (400, 112)])
def test_resize_tables_match_oracle(n_in, n_out):
    """vcagan_b200.preprocess.resize_coeffs (what the kernel is fed) vs oracle.pil_bilinear_tables: same integers."""
    from vcagan_b200.preprocess import resize_coeffs
    k, bounds = resize_coeffs(n_in, n_out)
    first, count, coef = O.pil_bilinear_tables(n_in, n_out)
    assert k.dtype == np.int32 and bounds.dtype == np.int32 and k.shape == coef.shape
    assert np.array_equal(bounds[:, 0], first) and np.array_equal(bounds[:, 1], count)
    assert np.array_equal(k.astype(np.int64), coef)
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all() and (bounds[:, 1] <= k.shape[1]).all()


def test_resize_tables_reproduce_pil_image_resize():
    Image = pytest.importorskip("PIL.Image")
    from vcagan_b200.preprocess import resize_coeffs
    rng = np.random.default_rng(3)
    for n_in in (136, 80, 231):
        row = rng.integers(0, 256, (1, n_in), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(row).resize((112, 1), Image.BILINEAR))[0].astype(np.int64)
        k, b = resize_coeffs(n_in, 112)
        got = np.array([np.clip(((1 << 21) + int(np.dot(row[0, b[o, 0]:b[o, 0] + b[o, 1]].astype(np.int64),
                                                      k[o, :b[o, 1]].astype(np.int64)))) >> 22, 0, 255) for o in range(112)])
        assert np.array_equal(got, ref), n_in


@pytest.mark.parametrize("fmax", [7500.0, 7600.0])
def test_mel_filterbank_matches_oracle(fmax):
    """The two independent restatements of librosa's Slaney basis (vectorised in the product, scalar loops in the
    oracle) agree to fp32 rounding."""
    from vcagan_b200.audio import mel_filterbank
    a = mel_filterbank(16000, 640, 80, 55.0, fmax)
    b = O.slaney_mel_basis(16000, 640, 80, 55.0, fmax)
    assert a.shape == b.shape == (80, 321) and a.dtype == np.float32
    assert np.abs(a.astype(np.float64) - b.astype(np.float64)).max() < 1e-9


def test_no_cpu_fallbacks():
    """Every public entry of the tail / preprocessing modules refuses CPU tensors instead of computing on the host."""
    import torch
    from vcagan_b200 import audio, preprocess
    with pytest.raises(RuntimeError):
        audio.deemphasize(torch.zeros(1, 16))
    with pytest.raises(RuntimeError):
        audio.griffin_lim(torch.zeros(1, 321, 4))
    with pytest.raises(RuntimeError):
        audio.lrs_denormalize_spec(torch.zeros(1, 321, 4))
    with pytest.raises(RuntimeError):
        preprocess.preprocess_clips(torch.zeros(1, 1, 8, 8, 3, dtype=torch.uint8))


def test_flat_group_slab_job_tables():
    """trainer.FlatGroup: tap-major gradient slabs are given to 2-D conv filters with more than one tap (Cin % 4 == 0) only; the
    job table of a flush over an element range [lo, hi) lists exactly the slab parameters whose flat offset lies inside it,
    with cumulative CTA offsets from vca_unslab_job_ctas and the largest tap count (host logic only: no kernel is launched)."""
    import torch
    import torch.nn as nn
    from vcagan_b200._lib import lib
    from vcagan_b200.trainer import FlatGroup
    a = nn.Sequential(nn.Conv2d(8, 16, 3), nn.BatchNorm2d(16), nn.Conv2d(16, 16, 1))       # 3x3 filter: slab; 1x1: none
    b = nn.Sequential(nn.Conv2d(16, 32, 5), nn.Linear(32, 4), nn.Conv2d(3, 8, 3))             # 5x5: slab; Cin = 3: none
    grp = FlatGroup([a, b])
    with_slab = [p for p in grp.params if hasattr(p, "_vca_slab")]
    assert [tuple(p.shape) for p in with_slab] == [(16, 8, 3, 3), (32, 16, 5, 5)]
    for p in with_slab:
        taps = p.shape[2] * p.shape[3]
        assert tuple(p._vca_slab.shape) == (taps, p.shape[0], p.shape[1]) and p._vca_slab.numel() == p.numel()
        assert p._vca_slab.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0
    tab, n, ctas, max_taps = grp._slab_table(0, grp.numel)
    assert n == 2 and max_taps == 25
    c0 = lib().query("vca_unslab_job_ctas", 16, 8, 9)
    assert tab[0, 3:7].tolist() == [16, 8, 9, 0] and tab[1, 3:7].tolist() == [32, 16, 25, c0]
    assert ctas == c0 + lib().query("vca_unslab_job_ctas", 32, 16, 25)
    cut = grp.offsets[len(list(a.parameters()))]                 # first parameter of module b
    t_lo = grp._slab_table(0, cut); t_hi = grp._slab_table(cut, grp.numel)
    assert t_lo[1] == 1 and t_hi[1] == 1 and t_lo[3] == 9 and t_hi[3] == 25 and t_hi[0][0, 6] == 0
    assert grp._slab_table(0, cut) is t_lo                      # cached
