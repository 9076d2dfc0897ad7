"""Test-time path of the reference (test.py:126-143): visual front-end + generator on the clip and on its
horizontally flipped copy (flip TTA, mels averaged BEFORE the Postnet), Postnet, then `inverse_spec` = Griffin-Lim on the
linear spectrogram, de-emphasis and clip (vid_aud_grid.py:212-224).  Everything stays on the device; nothing is copied to
the host until the waveform is returned."""
import torch

from . import audio


_side_streams = {}


def _front_and_generator(v_front, gen, vid, lens):
    """v_front -> gen (test.py:128-131) with the sentence GRU (4 x T strictly sequential steps on a few SMs: 4.9 ms for 128
    clips with most of the GPU idle) on a side stream underneath the generator's first six blocks, which need only the
    phoneme features -- the schedule Trainer._phase_d uses for training."""
    if not (hasattr(v_front, "features") and hasattr(gen, "stem")):
        phon, sent = v_front(vid)
        return gen(sent, phon, lens)[2]
    phon = v_front.features(vid)
    cur = torch.cuda.current_stream()
    side = _side_streams.get(vid.device)
    if side is None:
        side = _side_streams[vid.device] = torch.cuda.Stream(device=vid.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        sent = v_front.sentence(phon)
    h = gen.stem(phon)
    cur.wait_stream(side)
    sent.record_stream(cur)
    return gen.tail(sent, h, lens)[2]


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
        g = _front_and_generator(v_front, gen, torch.cat([vid, vid.flip(4)], 0), torch.cat([lens, lens], 0))
        g3 = (g[:B] + g[B:]) / 2.0
    else:
        g3 = _front_and_generator(v_front, gen, vid, vid_len)
    gs = post(g3)                                         # test.py:141
    spec = gs if mel_len is None else gs[..., :int(mel_len)]      # test.py:143 slices the whole batch to mel_len[0]
    mag = spec.squeeze(1).contiguous().float()
    if lrs:
        mag = audio.lrs_denormalize_spec(mag)
    wav_gl = audio.griffin_lim(mag, None, n_iters, init_angles=init_angles)            # vid_aud_grid.py:212-217
    return dict(mel=g3, spec=gs, wav=audio.deemphasize(wav_gl), wav_gl=wav_gl)          # :218-223


def save_eval_outputs(out_dir, f_names, mel, spec, mel_len, wav=None, sample_rate=16000):
    """What test.py:145-159 leaves on disk for the downstream ASR scorers (ASR_model/*/test.py read these files):
    `<out_dir>/spec_mel/<subject>/<file>.npz` with arrays `mel` = g3[b, :, :, :mel_len[b]] (1,80,L) and `spec` =
    gs[b, :, :, :mel_len[b]] (1,321,L), and -- when `wav` (B, samples) is given -- `<out_dir>/wav/<subject>/<file>.wav`
    as 16-bit PCM (the reference writes it with soundfile, subtype PCM_16; the stdlib `wave` module produces the same
    container).  f_names follow the dataset's 'subject/video/file' form (vid_aud_grid.py f_name).  Returns the paths."""
    import os
    import wave

    import numpy as np
    paths = []
    mel_c, spec_c = mel.detach().float().cpu().numpy(), spec.detach().float().cpu().numpy()
    wav_c = None if wav is None else (wav.detach().float().cpu().numpy() if torch.is_tensor(wav) else np.asarray(wav))
    for b, name in enumerate(f_names):
        sub_name, _, file_name = name.split('/')
        L = int(mel_len[b])
        d = os.path.join(out_dir, "spec_mel", sub_name)
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, file_name + ".npz")
        np.savez(p, mel=mel_c[b, :, :, :L], spec=spec_c[b, :, :, :L])
        paths.append(p)
        if wav_c is not None:
            d = os.path.join(out_dir, "wav", sub_name)
            os.makedirs(d, exist_ok=True)
            pcm = (np.clip(wav_c[b], -1.0, 1.0) * 32767.0).round().astype("<i2")
            with wave.open(os.path.join(d, file_name + ".wav"), "wb") as f:
                f.setnchannels(1); f.setsampwidth(2); f.setframerate(sample_rate)
                f.writeframes(pcm.tobytes())
    return paths
