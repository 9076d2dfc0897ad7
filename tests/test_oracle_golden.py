"""Pin the CPU oracle (oracle/vca_oracle.py) to vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json, os
import numpy as np
import pytest
import torch
from conftest import make_state, golden_inputs, rel_l2, GOLD
from oracle import vca_oracle as O

TOL = 2e-5  # same math, different op order (functional vs nn.Module) on CPU fp32


def test_known_answers(golden):
    assert abs(float(O.gan_loss(torch.zeros(4, 1), True)) - np.log(2)) < 1e-6
    assert abs(float(golden["known_gan_loss0"]) - np.log(2)) < 1e-6
    assert [O.final_length(t) for t in (40, 50, 75, 160, 250)] == [10, 12, 18, 40, 62]
    assert list(golden["known_final_length"]) == [10, 12, 18, 40, 62]


def test_eval_forward_matches_reference(golden, state_spec):
    vid, mel, spec, noise = golden_inputs()
    sds = {m: make_state(state_spec, m) for m in O.MODULES}
    with torch.no_grad():
        phon, sent = O.visual_front(sds["v_front"], vid, False)
        assert rel_l2(phon, golden["eval_phon"]) < TOL
        assert rel_l2(sent, golden["eval_sent"]) < TOL
        g1, g2, g3 = O.decoder(sds["gen"], sent, phon, [20, 13], noise, False)
        for a, k in ((g1, "eval_g1"), (g2, "eval_g2"), (g3, "eval_g3")):
            assert rel_l2(a, golden[k]) < TOL, k
        gs = O.postnet(sds["post"], g3, False)
        assert rel_l2(gs, golden["eval_gs"]) < TOL
        mel1, mel2 = O.bilinear_half(mel, 0.25), O.bilinear_half(mel, 0.5)
        for i, x in ((1, mel1), (2, mel2), (3, mel)):
            u, c = O.discriminator(sds[f"dis{i}"], x, sent, 20)
            assert rel_l2(u, golden[f"eval_d{i}_u"]) < TOL
            assert rel_l2(c, golden[f"eval_d{i}_c"]) < TOL
        assert rel_l2(O.sync_discriminator(sds["s_dis"], phon, mel, False, False), golden["eval_sync_nce"]) < TOL
        assert rel_l2(O.sync_discriminator(sds["s_dis"], phon, g3, True, False), golden["eval_sync_cos"]) < TOL


def test_masked_keys_do_not_matter(state_spec):
    """keys at positions >= len must not influence the attention output."""
    sd = make_state(state_spec, "gen")
    g = torch.Generator().manual_seed(5)
    ph = torch.randn(2, 20, 512, generator=g); feat = torch.randn(2, 128, 20, 20, generator=g)
    a = O.av_attention(sd, "att1", ph, feat, [20, 13])
    ph2 = ph.clone(); ph2[1, 13:] += 10.0
    b = O.av_attention(sd, "att1", ph2, feat, [20, 13])
    assert torch.allclose(a[1], b[1], atol=1e-6) and torch.allclose(a[0], b[0])


def test_train_step_matches_reference(golden, state_spec):
    vid, mel, spec, noise = golden_inputs()
    sds = {m: make_state(state_spec, m, requires_grad=True) for m in O.MODULES}
    par = lambda ms: [{"params": [p for p in sds[m].values() if p.requires_grad]} for m in ms]
    g_opt = torch.optim.Adam(par(("v_front", "gen", "post")), lr=1e-4, weight_decay=1e-5, amsgrad=True)
    d_opt = torch.optim.Adam(par(("dis1", "dis2", "dis3", "s_dis")), lr=1e-4, weight_decay=1e-5, amsgrad=True)
    out = O.train_step_with_adam(sds, dict(mel=mel, spec=spec, vid=vid, vid_len=[20, 13]), noise, g_opt, d_opt)
    for k in ("dis_loss", "sync_loss", "real_loss", "fake_loss", "gen_loss", "g_sync", "recon"):
        ref = float(golden["step_" + k]); got = float(out[k])
        assert abs(got - ref) <= 2e-5 * max(1.0, abs(ref)), (k, got, ref)
    assert rel_l2(out["grad_pen"], golden["step_grad_pen"]) < 1e-4
    for k in ("g1", "g2", "g3", "gs"):
        assert rel_l2(out[k], golden["step_" + k]) < TOL, k
    assert rel_l2(out["r1_grads"][2], golden["step_r1_grad3"]) < 1e-4
    assert rel_l2(out["r1_grads"][0], golden["step_r1_grad1"]) < 1e-4
    names = json.load(open(os.path.join(GOLD, "grad_norm_names.json")))
    dn = torch.tensor([out["d_grad_norms"][n.split(".", 1)[0]][n.split(".", 1)[1]] for n in names["d"]])
    assert rel_l2(dn, golden["step_d_grad_norms"]) < 1e-4
    gn = torch.tensor([out["g_grad_norms"][n.split(".", 1)[0]][n.split(".", 1)[1]] for n in names["g"]])
    assert rel_l2(gn, golden["step_g_grad_norms"]) < 1e-4
    vfd = torch.tensor([float(out["vf_grad_after_d"][n.split(".", 1)[1]].norm()) for n in names["vf_d"]])
    assert rel_l2(vfd, golden["step_vf_d_grad_norms"]) < 1e-4
    cn = json.load(open(os.path.join(GOLD, "checksum_names.json")))
    chk = torch.tensor([[float(sds[n.split(".", 1)[0]][n.split(".", 1)[1]].double().sum()),
                         float(sds[n.split(".", 1)[0]][n.split(".", 1)[1]].double().abs().sum())] for n in cn["params"]],
                       dtype=torch.float64)
    ref = torch.from_numpy(golden["step_param_checksums"])
    assert float((chk[:, 1] - ref[:, 1]).abs().max() / ref[:, 1].abs().max()) < 1e-6
    bs = torch.tensor([float(sds[n.split(".", 1)[0]][n.split(".", 1)[1]].double().sum()) for n in cn["buffers"]],
                      dtype=torch.float64)
    assert rel_l2(bs, golden["step_buffer_sums"]) < 1e-5


def test_stft_griffin_lim_matches_reference(golden):
    fwd, inv = O.stft_bases()
    g = torch.Generator().manual_seed(77)
    sig = torch.randn(2, 160 * 11, generator=g) * 0.1
    mag, ph = O.stft_transform(sig, fwd)
    assert rel_l2(mag, golden["stft_mag"]) < 1e-6
    assert rel_l2(torch.cos(ph), np.cos(golden["stft_phase"])) < 1e-5
    rec = O.stft_inverse(mag, ph, inv)
    assert rel_l2(rec, golden["stft_rec"]) < 1e-6
    assert float((rec.squeeze(1) - sig).abs().max()) < 1e-5  # SURVEY section 4 item 5 round trip
    assert np.allclose(O.window_sumsquare(12), golden["window_sumsquare_12"], atol=1e-7)
    wav = O.griffin_lim(torch.from_numpy(golden["gl_mag"]), torch.from_numpy(golden["gl_init_phase"]), 8)
    assert rel_l2(wav, golden["gl_wav"]) < 1e-5
    # torch.stft cross-check (hann periodic, center, reflect)
    ts = torch.stft(sig, 640, 160, 640, torch.hann_window(640, periodic=True), center=True, pad_mode="reflect",
                    return_complex=True)
    assert rel_l2(mag, ts.abs()) < 1e-5


# ---- waveform tail / mel front -----------------------------------------------------
def test_deemphasis_matches_reference_and_scipy(golden_tail):
    x = golden_tail["deemph_in"]
    y = O.deemphasize_clip(x)
    assert np.abs(y - np.clip(golden_tail["deemph_out"], -1, 1)).max() < 1e-12
    signal = pytest.importorskip("scipy.signal")
    ref = np.stack([signal.lfilter([1], [1, -0.97], w) for w in x])     # vid_aud_grid.py:230-232
    assert np.abs(y - np.clip(ref, -1, 1)).max() < 1e-12


def test_mel_basis_properties():
    """librosa is absent, so the Slaney basis is checked through the properties its definition gives: shape, triangular
    non-negative filters with a single peak, area normalisation 2 / (f[m+2] - f[m]), support inside [fmin, fmax]."""
    for fmax in (7500.0, 7600.0):                    # GRID vid_aud_grid.py:37, LRS vid_aud_lrs2.py
        b = O.slaney_mel_basis(16000, 640, 80, 55.0, fmax).astype(np.float64)
        assert b.shape == (80, 321) and (b >= 0).all()
        freqs = np.arange(321) * 25.0
        assert b[:, freqs < 55.0].sum() == 0 and b[:, freqs > fmax].sum() == 0
        for m in range(80):
            nz = np.nonzero(b[m])[0]
            assert len(nz) > 0 and (np.diff(nz) == 1).all()
            pk = nz[np.argmax(b[m, nz])]
            assert (np.diff(b[m, nz[0]:pk + 1]) >= -1e-12).all() and (np.diff(b[m, pk:nz[-1] + 1]) <= 1e-12).all()
        # a triangle of height h over [l, r] integrates to h (r - l) / 2 = 1 for every filter (area normalisation);
        # sampled every 25 Hz the sum * 25 approaches 1 for the wide high filters
        assert np.abs(b[40:].sum(1) * 25.0 - 1.0).max() < 0.05


def test_tail_restatement_matches_reference(golden_tail):
    gt = golden_tail
    basis = O.slaney_mel_basis()
    assert np.array_equal(basis, gt["mel_basis"])
    assert rel_l2(O.mel_to_spec(torch.from_numpy(gt["mel"]), basis), gt["mel_to_spec"]) < 1e-6
    ph = torch.from_numpy(gt["grid_phase"])
    wav = O.deemphasize_clip(O.griffin_lim(torch.from_numpy(gt["grid_spec"]).squeeze(1), ph, 60).numpy())
    assert rel_l2(wav, gt["grid_inverse_spec"]) < 1e-4                 # vid_aud_grid.py:212-224
    wav = O.deemphasize_clip(O.griffin_lim(O.mel_to_spec(torch.from_numpy(gt["mel"]), basis), ph, 60).numpy())
    assert rel_l2(wav, gt["grid_inverse_mel"]) < 1e-4                  # vid_aud_grid.py:190-210
    mag = O.lrs_denormalize_spec(torch.from_numpy(gt["lrs_spec"])).squeeze(1)
    wav = O.deemphasize_clip(O.griffin_lim(mag, ph, 60).numpy())
    assert rel_l2(wav, gt["lrs_inverse_spec"]) < 1e-4                  # vid_aud_lrs2.py:257-272
    mel, mags = O.mel_spectrogram(torch.from_numpy(gt["melspec_in"]), basis)
    assert rel_l2(mags, gt["melspec_mag"]) < 1e-6 and rel_l2(mel, gt["melspec_out"]) < 1e-5


# ---- clip preprocessing of the loader ---------------------------------------------
def _lrs_boxes(centres, s):
    c = np.asarray(centres).reshape(-1, 2)
    return np.stack([c[:, 0] - 40 + s, c[:, 1] - 40 + s, c[:, 0] + 40 + s, c[:, 1] + 40 + s], 1)   # vid_aud_lrs2.py:93-98


def preproc_cases(gp):
    """(name, frames, boxes, max_t, flip, erase) for every vector of golden_preproc.npz."""
    from conftest import synthetic_frames
    fg, fl = synthetic_frames(11, 3, 288, 360), synthetic_frames(12, 4, 160, 160)
    cases = [("grid_plain", fg, np.array([59, 95, 195, 231]), 5, False, None)]
    for seed in (3, 4, 10):
        flip, xs, ys = gp[f"grid_aug{seed}_draws"]
        cases.append((f"grid_aug{seed}", fg, np.array([59, 95, 195, 231]), 5, bool(flip), (int(xs), int(ys))))
    cases.append(("lrs_plain", fl, _lrs_boxes(gp["lrs_centres"], 0), 6, False, None))
    for seed in (1, 2):
        s, flip = gp[f"lrs_aug{seed}_draws"]
        cases.append((f"lrs_aug{seed}", fl, _lrs_boxes(gp["lrs_centres"], int(s)), 6, bool(flip), None))
    return cases


def test_preprocess_restatement_is_bit_exact(golden_preproc):
    flips = set()
    for name, frames, boxes, max_t, flip, erase in preproc_cases(golden_preproc):
        got = O.preprocess_clip(frames, boxes, max_t, flip, erase)
        assert torch.equal(got, torch.from_numpy(golden_preproc[name])), name
        flips.add(flip)
    assert flips == {False, True}          # the seeds cover both branches of the flip


def test_pil_tables_against_live_pil():
    """Where PIL is installed, pin the coefficient restatement to the library itself (an impulse through Image.resize
    reads the fixed-point kernel back) for the two shrink / enlarge factors of the path and a few others."""
    Image = pytest.importorskip("PIL.Image")
    for n_in, n_out in ((136, 112), (80, 112), (112, 112), (300, 112), (57, 112)):
        first, count, coef = O.pil_bilinear_tables(n_in, n_out)
        assert (coef.sum(1) > 0).all() and abs(coef.sum(1) - (1 << 22)).max() <= 3
        for pos in (0, n_in // 3, n_in - 1):
            line = np.zeros((1, n_in), np.uint8); line[0, pos] = 255
            ref = np.asarray(Image.fromarray(line).resize((n_out, 1), Image.BILINEAR))[0]
            mine = np.zeros(n_out, np.int64)
            for o in range(n_out):
                j = pos - first[o]
                if 0 <= j < count[o]:
                    mine[o] = np.clip(((1 << 21) + 255 * coef[o, j]) >> 22, 0, 255)
            assert np.array_equal(mine, ref), (n_in, n_out, pos)
