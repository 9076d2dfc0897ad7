"""Test-time path of the reference (test.py:126-143): visual front-end + generator on the clip and on its
horizontally flipped copy (flip TTA, mels averaged BEFORE the Postnet), Postnet, then `inverse_spec` = Griffin-Lim on the
linear spectrogram, de-emphasis and clip (vid_aud_grid.py:212-224).  Everything stays on the device; nothing is copied to
the host until the waveform is returned."""
import torch

from . import audio


@torch.no_grad()
def synthesize(v_front, gen, post, vid, vid_len, n_iters=60, tta=True, mel_len=None, init_angles=None, lrs=False):
    """vid (B,1,T,112,112) on the GPU -> dict(mel g3 (B,1,80,4T), spec gs (B,1,321,4T), wav (B, 160*(L-1)) de-emphasised
    and clipped to [-1,1] as test.py:143 saves it, wav_gl = the raw Griffin-Lim signal).  lrs=True applies the LRS
    spectrogram de-normalisation of vid_aud_lrs2.py:261-263 first (test_LRS.py:161)."""
    for m in (v_front, gen, post):
        m.eval()
    if tta:
        # test.py:134-140 runs the clip and its mirror image one after the other.  In eval mode nothing couples the
        # samples of a batch (BatchNorm uses its running statistics), so both go through as ONE batch of 2B: the same
        # per-sample arithmetic, half the launches, twice the rows per GEMM tile wave.
        B = vid.shape[0]
        lens = torch.as_tensor(vid_len).reshape(-1)
        phon, sent = v_front(torch.cat([vid, vid.flip(4)], 0))
        g = gen(sent, phon, torch.cat([lens, lens], 0))[2]
        g3 = (g[:B] + g[B:]) / 2.0
    else:
        phon, sent = v_front(vid)
        g3 = gen(sent, phon, vid_len)[2]
    gs = post(g3)                                         # test.py:141
    spec = gs if mel_len is None else gs[..., :int(mel_len)]      # test.py:143 slices the whole batch to mel_len[0]
    mag = spec.squeeze(1).contiguous().float()
    if lrs:
        mag = audio.lrs_denormalize_spec(mag)
    wav_gl = audio.griffin_lim(mag, None, n_iters, init_angles=init_angles)            # vid_aud_grid.py:212-217
    return dict(mel=g3, spec=gs, wav=audio.deemphasize(wav_gl), wav_gl=wav_gl)          # :218-223
