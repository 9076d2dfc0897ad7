"""CPU oracle for the VCA-GAN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *functional restatement* (plain torch fp32/fp64 ops over a flat
``{name: tensor}`` dict that uses the reference's ``state_dict`` key names) of
the algorithm in the reference's ``src/models`` and ``src/data/{stft,
audio_processing}.py``.  It is the checker used by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  Nothing on the product path may import it: the product
path is the CUDA library under ``visual-context-attentional-gan_b200/``.

Parity pin: the reference ships no golden vectors, so the
pin is ``tests/golden/*.npz`` -- outputs of the *unmodified reference modules*
imported from /root/reference by ``tests/golden/make_golden.py`` (models, one
full G+D step, STFT / Griffin-Lim), ``make_golden_tail.py`` (the dataset
classes' inverse_spec / inverse_mel / deemphasize, TacotronSTFT.mel_spectrogram)
and ``make_golden_preproc.py`` (build_tensor: the PIL / torchvision clip
preprocessing), all committed; ``tests/test_oracle_golden.py`` holds this
restatement to those vectors.  UNPINNED: one matrix, the librosa mel basis
(librosa is not installed; see ``slaney_mel_basis``).

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
LRELU = 0.2
BN_EPS = 1e-5
BN_MOM = 0.1
INV_SQRT2 = 1.0 / math.sqrt(2.0)
LOG1E5 = math.log(1e-5)  # src/data/vid_aud_grid.py:21 (log1e5)


# --------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------
def _bn(sd: SD, pre: str, x: torch.Tensor, train: bool) -> torch.Tensor:
    """nn.BatchNorm{1,2,3}d with torch defaults (eps 1e-5, momentum 0.1);
    running stats in ``sd`` are updated in place when ``train``."""
    y = F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"],
                     sd[pre + ".weight"], sd[pre + ".bias"], train, BN_MOM, BN_EPS)
    if train and (pre + ".num_batches_tracked") in sd:
        sd[pre + ".num_batches_tracked"] += 1
    return y


def _prelu(sd: SD, key: str, x: torch.Tensor) -> torch.Tensor:
    return F.prelu(x, sd[key])


def _lrelu(x: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(x, LRELU)


def final_length(t: int) -> int:
    """src/models/generator.py:368-371"""
    return (t // 2) // 2


def gan_loss(logits: torch.Tensor, real: Optional[bool] = None) -> torch.Tensor:
    """src/models/generator.py:363-366 -- softplus(-x) for label True, softplus(x) otherwise."""
    return F.softplus(-logits if real else logits).mean()


def denormalize(mel: torch.Tensor) -> torch.Tensor:
    """src/data/vid_aud_grid.py:238-240"""
    return (mel + 1.0) * (-LOG1E5 / 2.0) + LOG1E5


# --------------------------------------------------------------------------
# visual front-end  (src/models/visual_front.py, src/models/resnet.py)
# --------------------------------------------------------------------------
def _basic_block(sd: SD, pre: str, x: torch.Tensor, stride: int, train: bool, prelu: bool) -> torch.Tensor:
    """src/models/resnet.py:25-66"""
    y = F.conv2d(x, sd[pre + ".conv1.weight"], None, stride, 1)
    y = _bn(sd, pre + ".bn1", y, train)
    y = _prelu(sd, pre + ".relu1.weight", y) if prelu else F.relu(y)
    y = F.conv2d(y, sd[pre + ".conv2.weight"], None, 1, 1)
    y = _bn(sd, pre + ".bn2", y, train)
    if (pre + ".downsample.0.weight") in sd:
        x = F.conv2d(x, sd[pre + ".downsample.0.weight"], None, stride, 0)
        x = _bn(sd, pre + ".downsample.1", x, train)
    y = y + x
    return _prelu(sd, pre + ".relu2.weight", y) if prelu else F.relu(y)


def gru_bidir_2layer(sd: SD, pre: str, x: torch.Tensor,
                     drop_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.GRU(512,512,2,bidirectional) as used at src/models/visual_front.py:20,33-34.
    x: (T,B,512) -> (T,B,1024).  Gate order r,z,n; b_hn inside the r-product
    (SURVEY.md appendix D).  ``drop_mask`` (T,B,1024), already scaled by 1/(1-p),
    is the inter-layer dropout; None = no dropout."""
    T, B, _ = x.shape
    inp = x
    for layer in range(2):
        outs = []
        for rev in (False, True):
            sfx = f"_l{layer}" + ("_reverse" if rev else "")
            w_ih, w_hh = sd[f"{pre}.weight_ih{sfx}"], sd[f"{pre}.weight_hh{sfx}"]
            b_ih, b_hh = sd[f"{pre}.bias_ih{sfx}"], sd[f"{pre}.bias_hh{sfx}"]
            H = w_hh.shape[1]
            gi_all = F.linear(inp, w_ih, b_ih)  # (T,B,3H)
            h = x.new_zeros(B, H)
            seq = [None] * T
            order = range(T - 1, -1, -1) if rev else range(T)
            for t in order:
                gh = F.linear(h, w_hh, b_hh)
                gi = gi_all[t]
                r = torch.sigmoid(gi[:, :H] + gh[:, :H])
                z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                h = (1.0 - z) * n + z * h
                seq[t] = h
            outs.append(torch.stack(seq, 0))
        inp = torch.cat(outs, 2)
        if layer == 0 and drop_mask is not None:
            inp = inp * drop_mask
    return inp


def visual_front(sd: SD, vid: torch.Tensor, train: bool,
                 drop_feat: Optional[torch.Tensor] = None,
                 drop_gru: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """src/models/visual_front.py:23-37.  vid (B,1,T,112,112) -> phons (B,T,512), sentence (B,512,T).
    ``drop_feat`` (B*T,512) / ``drop_gru`` (T,B,1024) are pre-scaled dropout masks
    (None = dropout off), injected so both sides of a parity test see the same mask."""
    x = F.conv3d(vid, sd["frontend.0.weight"], None, (1, 2, 2), (2, 3, 3))
    x = _bn(sd, "frontend.1", x, train)
    x = _prelu(sd, "frontend.2.weight", x)
    x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    B, C, T, H, W = x.shape
    x = x.transpose(1, 2).reshape(B * T, C, H, W)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = _basic_block(sd, f"resnet.layer{li}.0", x, stride, train, True)
        x = _basic_block(sd, f"resnet.layer{li}.1", x, 1, train, True)
    x = F.avg_pool2d(x, 4).flatten(1)  # (B*T,512)
    if drop_feat is not None:
        x = x * drop_feat
    phons_tb = x.view(B, T, -1).permute(1, 0, 2).contiguous()  # (T,B,512)
    s = gru_bidir_2layer(sd, "sentence_encoder", phons_tb, drop_gru)
    s = F.linear(s, sd["fc.weight"], sd["fc.bias"]).permute(1, 2, 0).contiguous()
    return phons_tb.permute(1, 0, 2), s


# --------------------------------------------------------------------------
# generator  (src/models/generator.py:94-265)
# --------------------------------------------------------------------------
def _gen_res_blk(sd: SD, pre: str, x: torch.Tensor, up: bool, train: bool) -> torch.Tensor:
    """GenResBlk, src/models/generator.py:94-131 (BN first; shortcut upsamples then 1x1)."""
    r = _lrelu(_bn(sd, pre + ".norm1", x, train))
    if up:
        r = F.interpolate(r, scale_factor=2, mode="nearest")
    r = F.conv2d(r, sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], 1, 2)
    r = _lrelu(_bn(sd, pre + ".norm2", r, train))
    r = F.conv2d(r, sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], 1, 2)
    s = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    if (pre + ".conv1x1.weight") in sd:
        s = F.conv2d(s, sd[pre + ".conv1x1.weight"])
    return (r + s) * INV_SQRT2


def av_attention(sd: SD, pre: str, ph: torch.Tensor, g: torch.Tensor, lens: Sequence[int]) -> torch.Tensor:
    """AVAttention.forward, src/models/generator.py:154-171.  ph (B,S,512), g (B,C,F,T) -> (B,C',F,T)."""
    B, C, Fq, T = g.shape
    k = F.linear(ph, sd[pre + ".k.weight"], sd[pre + ".k.bias"])  # B,S,256
    v = F.linear(ph, sd[pre + ".v.weight"], sd[pre + ".v.bias"])
    q = F.linear(g.reshape(B, C * Fq, T).transpose(1, 2), sd[pre + ".q.weight"], sd[pre + ".q.bias"])
    d = k.shape[-1]
    att = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(d)  # B,T,S
    S = att.shape[2]
    key_idx = torch.arange(S).view(1, 1, S)
    lens_t = torch.as_tensor(list(int(l) for l in lens)).view(B, 1, 1)
    att = att.masked_fill(key_idx >= lens_t, float("-inf"))
    att = torch.softmax(att, 2)
    o = F.linear(torch.bmm(att, v), sd[pre + ".mel.weight"], sd[pre + ".mel.bias"])  # B,T,1280
    return o.view(B, T, Fq, -1).permute(0, 3, 2, 1)


def _to_mel(sd: SD, pre: str, x: torch.Tensor, train: bool) -> torch.Tensor:
    """to_mel{1,2,3}, src/models/generator.py:208-225"""
    x = _lrelu(_bn(sd, pre + ".0", x, train))
    return torch.tanh(F.conv2d(x, sd[pre + ".2.weight"], sd[pre + ".2.bias"]))


def decoder(sd: SD, sent: torch.Tensor, phon: torch.Tensor, lens: Sequence[int],
            noise: torch.Tensor, train: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Decoder.forward, src/models/generator.py:245-265.  ``noise`` (B,128,20,T) replaces the
    host-side torch.randn of line 248."""
    s = sent.transpose(1, 2)  # B,T,512
    x = phon.transpose(1, 2).unsqueeze(2).expand(-1, -1, 20, -1)
    x = torch.cat([x, noise], 1)
    for i in range(3):
        x = _gen_res_blk(sd, f"decode.{i}", x, False, train)
    for i in range(3):
        x = _gen_res_blk(sd, f"g1.{i}", x, False, train)
    f1 = x
    c1 = av_attention(sd, "att1", s, f1, lens)
    x = F.conv2d(torch.cat([x, c1], 1), sd["attconv1.weight"], sd["attconv1.bias"], 1, 2)
    for i in range(3):
        x = _gen_res_blk(sd, f"g2.{i}", x, i == 0, train)
    f2 = x
    c2 = av_attention(sd, "att2", s, f2, lens)
    x = F.conv2d(torch.cat([x, c2], 1), sd["attconv2.weight"], sd["attconv2.bias"], 1, 2)
    for i in range(3):
        x = _gen_res_blk(sd, f"g3.{i}", x, i == 0, train)
    return _to_mel(sd, "to_mel1", f1, train), _to_mel(sd, "to_mel2", f2, train), _to_mel(sd, "to_mel3", x, train)


def _res_blk1d(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """ResBlk1D (normalize=False, downsample=False), src/models/generator.py:8-49"""
    r = F.conv1d(_lrelu(x), sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], 1, 2)
    r = F.conv1d(_lrelu(r), sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], 1, 2)
    s = F.conv1d(x, sd[pre + ".conv1x1.weight"]) if (pre + ".conv1x1.weight") in sd else x
    return (s + r) * INV_SQRT2


def postnet(sd: SD, mel: torch.Tensor, train: bool) -> torch.Tensor:
    """Postnet.forward, src/models/generator.py:173-192.  (B,1,80,L) -> (B,1,321,L)"""
    x = mel.squeeze(1)
    x = F.conv1d(x, sd["postnet.0.weight"], sd["postnet.0.bias"], 1, 3)
    x = _lrelu(_bn(sd, "postnet.1", x, train))
    for i in (3, 4, 5):
        x = _res_blk1d(sd, f"postnet.{i}", x)
    return F.conv1d(x, sd["postnet.6.weight"]).unsqueeze(1)


# --------------------------------------------------------------------------
# discriminators  (src/models/generator.py:51-92, 267-361)
# --------------------------------------------------------------------------
def _res_blk_down(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """ResBlk(downsample=True, normalize=False), src/models/generator.py:51-92"""
    r = F.conv2d(_lrelu(x), sd[pre + ".conv1.weight"], sd[pre + ".conv1.bias"], 1, 2)
    r = F.avg_pool2d(r, 2)
    r = F.conv2d(_lrelu(r), sd[pre + ".conv2.weight"], sd[pre + ".conv2.bias"], 1, 2)
    s = F.conv2d(x, sd[pre + ".conv1x1.weight"]) if (pre + ".conv1x1.weight") in sd else x
    s = F.avg_pool2d(s, 2)
    return (s + r) * INV_SQRT2


def discriminator(sd: SD, x: torch.Tensor, c: torch.Tensor, vid_max_length: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Discriminator.forward, src/models/generator.py:306-317; number of ResBlks read off the keys."""
    n_blk = len({k.split(".")[1] for k in sd if k.startswith("main.")}) - 1
    f_len = final_length(vid_max_length)
    cm = c.mean(2)[:, :, None, None].expand(-1, -1, 5, f_len)
    h = F.conv2d(x, sd["main.0.weight"], sd["main.0.bias"], 1, 2)
    for i in range(1, n_blk + 1):
        h = _res_blk_down(sd, f"main.{i}", h)
    u = F.conv2d(_lrelu(h), sd["uncond.1.weight"], sd["uncond.1.bias"])
    u = F.linear(_lrelu(u).mean([2, 3]), sd["uncond.4.weight"], sd["uncond.4.bias"])
    k = F.conv2d(_lrelu(torch.cat([h, cm], 1)), sd["cond.1.weight"], sd["cond.1.bias"], 1, 2)
    k = F.conv2d(_lrelu(k), sd["cond.3.weight"], sd["cond.3.bias"])
    k = F.linear(_lrelu(k).mean([2, 3]), sd["cond.6.weight"], sd["cond.6.bias"])
    return u.view(u.size(0), -1), k.view(k.size(0), -1)


def sync_discriminator(sd: SD, v_feat: torch.Tensor, aud: torch.Tensor, gen: bool, train: bool,
                       temp: float = 1.0) -> torch.Tensor:
    """sync_Discriminator.forward, src/models/generator.py:339-361.  -> (B,)"""
    a = F.conv2d(aud, sd["frontend.0.weight"], sd["frontend.0.bias"], 2, 1)
    a = _prelu(sd, "frontend.2.weight", _bn(sd, "frontend.1", a, train))
    a = F.conv2d(a, sd["frontend.3.weight"], sd["frontend.3.bias"], 2, 1)
    a = _prelu(sd, "frontend.5.weight", _bn(sd, "frontend.4", a, train))
    a = _basic_block(sd, "Res_block.0", a, 1, train, False)
    b, c, f, t = a.shape
    a = F.linear(a.reshape(b, c * f, t).transpose(1, 2), sd["Linear.weight"], sd["Linear.bias"])
    if gen:
        return 5.0 - F.cosine_similarity(v_feat, a, 2).abs().mean(1)
    vn, an = F.normalize(v_feat, dim=2), F.normalize(a, dim=2)
    sim = torch.bmm(vn, an.transpose(1, 2)) / temp
    va = torch.diagonal(F.log_softmax(sim, 2), dim1=-2, dim2=-1).mean(1)
    av = torch.diagonal(F.log_softmax(sim, 1), dim1=-2, dim2=-1).mean(1)
    return -0.5 * (va + av)


# --------------------------------------------------------------------------
# STFT / Griffin-Lim  (src/data/stft.py, src/data/audio_processing.py)
# --------------------------------------------------------------------------
def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), src/data/stft.py:59"""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_bases(n_fft: int = 640, hop: int = 160) -> Tuple[torch.Tensor, torch.Tensor]:
    """src/data/stft.py:45-68: (2*cutoff,1,n_fft) windowed forward basis and pinv-based inverse basis."""
    fb = np.fft.fft(np.eye(n_fft))
    cut = n_fft // 2 + 1
    fb = np.vstack([np.real(fb[:cut]), np.imag(fb[:cut])])
    fwd = torch.FloatTensor(fb[:, None, :])
    inv = torch.FloatTensor(np.linalg.pinv((n_fft / hop) * fb).T[:, None, :])
    w = torch.from_numpy(hann_periodic(n_fft)).float()
    return (fwd * w).float(), (inv * w).float()


def window_sumsquare(n_frames: int, n_fft: int = 640, hop: int = 160) -> np.ndarray:
    """src/data/audio_processing.py:7-48 (win_length == n_fft, norm None)."""
    n = n_fft + hop * (n_frames - 1)
    env = np.zeros(n, dtype=np.float32)
    wsq = hann_periodic(n_fft) ** 2
    for i in range(n_frames):
        s = i * hop
        env[s:min(n, s + n_fft)] += wsq[:max(0, min(n_fft, n - s))]
    return env


def stft_transform(x: torch.Tensor, fwd: torch.Tensor, n_fft: int = 640, hop: int = 160):
    """STFT.transform, src/data/stft.py:70-98.  x (B,L) -> magnitude, phase (B,n_fft/2+1,frames)."""
    xp = F.pad(x[:, None, None, :], (n_fft // 2, n_fft // 2, 0, 0), mode="reflect").squeeze(1)
    ft = F.conv1d(xp, fwd, stride=hop)
    cut = n_fft // 2 + 1
    re, im = ft[:, :cut], ft[:, cut:]
    return torch.sqrt(re * re + im * im), torch.atan2(im, re)


def stft_inverse(mag: torch.Tensor, phase: torch.Tensor, inv: torch.Tensor,
                 n_fft: int = 640, hop: int = 160) -> torch.Tensor:
    """STFT.inverse, src/data/stft.py:100-129 -> (B,1,hop*(frames-1))."""
    z = torch.cat([mag * torch.cos(phase), mag * torch.sin(phase)], 1)
    y = F.conv_transpose1d(z, inv, stride=hop)
    ws = window_sumsquare(mag.size(-1), n_fft, hop)
    nz = torch.from_numpy(np.where(ws > np.finfo(np.float32).tiny)[0])
    wst = torch.from_numpy(ws)
    y[:, :, nz] = y[:, :, nz] / wst[nz]
    y = y * (float(n_fft) / hop)
    return y[:, :, n_fft // 2:-(n_fft // 2)]


def griffin_lim(mag: torch.Tensor, init_phase: torch.Tensor, n_iters: int = 60,
                n_fft: int = 640, hop: int = 160) -> torch.Tensor:
    """griffin_lim, src/data/audio_processing.py:51-68, with the numpy-RNG initial phase of
    lines 59-62 injected as ``init_phase`` (B,n_fft/2+1,frames).  -> (B, hop*(frames-1))."""
    fwd, inv = stft_bases(n_fft, hop)
    sig = stft_inverse(mag, init_phase, inv, n_fft, hop).squeeze(1)
    for _ in range(n_iters):
        _, ang = stft_transform(sig, fwd, n_fft, hop)
        sig = stft_inverse(mag, ang, inv, n_fft, hop).squeeze(1)
    return sig


# --------------------------------------------------------------------------
# waveform tail / mel front  (src/data/vid_aud_grid.py:190-240, 270-307;
# src/data/vid_aud_lrs2.py:257-296).  The mel basis comes from librosa
# (librosa.filters.mel, un-vendored, no pinned version; the reference calls the
# positional librosa<0.10 signature at vid_aud_grid.py:278) which is absent
# here: `slaney_mel_basis` restates its published algorithm -- PARITY UNPINNED
# for that one matrix.  The de-emphasis filter is pinned against
# scipy.signal.lfilter in tests/test_oracle_golden.py.
# --------------------------------------------------------------------------
def deemphasize_clip(wav: np.ndarray, coef: float = 0.97) -> np.ndarray:
    """signal.lfilter([1], [1, -coef], w) per waveform in float64, then np.clip(-1, 1)
    (vid_aud_grid.py:205-209, 230-232).  wav (B,L) -> float64 (B,L)."""
    x = np.asarray(wav, dtype=np.float64)
    y = np.empty_like(x)
    state = np.zeros(x.shape[0], dtype=np.float64)
    for n in range(x.shape[1]):
        state = x[:, n] + coef * state
        y[:, n] = state
    return np.clip(y, -1.0, 1.0)


def slaney_mel_basis(sr: int = 16000, n_fft: int = 640, n_mels: int = 80,
                     fmin: float = 55.0, fmax: float = 7500.0) -> np.ndarray:
    """librosa.filters.mel defaults (htk=False, norm='slaney') as used at vid_aud_grid.py:278."""
    def to_mel(hz: float) -> float:
        if hz < 1000.0:
            return 3.0 * hz / 200.0
        return 15.0 + 27.0 * math.log(hz / 1000.0) / math.log(6.4)

    def to_hz(mel: float) -> float:
        if mel < 15.0:
            return 200.0 * mel / 3.0
        return 1000.0 * math.exp(math.log(6.4) * (mel - 15.0) / 27.0)

    lo, hi = to_mel(fmin), to_mel(fmax)
    edges = [to_hz(lo + (hi - lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    n_bins = n_fft // 2 + 1
    basis = np.zeros((n_mels, n_bins), dtype=np.float64)
    for m in range(n_mels):
        left, centre, right = edges[m], edges[m + 1], edges[m + 2]
        for k in range(n_bins):
            f = k * (sr / 2.0) / (n_bins - 1)
            tri = min((f - left) / (centre - left), (right - f) / (right - centre))
            basis[m, k] = max(0.0, tri) * 2.0 / (right - left)
    return basis.astype(np.float32)


def mel_to_spec(mel: torch.Tensor, basis: np.ndarray) -> torch.Tensor:
    """Front of inverse_mel, vid_aud_grid.py:194-200: (B,1,80,T) normalised mel -> (B,321,T)."""
    m = torch.exp(denormalize(mel))                      # :194-195 (spectral_de_normalize = exp, C = 1)
    m = m.transpose(2, 3).contiguous()                   # B,1,T,80
    s = torch.matmul(m, torch.from_numpy(basis))         # :198
    return s.transpose(2, 3).squeeze(1) * 1000.0         # :199-200


def lrs_denormalize_spec(spec: torch.Tensor) -> torch.Tensor:
    """vid_aud_lrs2.py:261-263: denormalize -> exp -> * 14."""
    return torch.exp(denormalize(spec)) * 14.0


def mel_spectrogram(y: torch.Tensor, basis: np.ndarray, n_fft: int = 640, hop: int = 160):
    """TacotronSTFT.mel_spectrogram, vid_aud_grid.py:291-307: (B,L) -> (log-mel (B,80,frames), magnitudes)."""
    fwd, _ = stft_bases(n_fft, hop)
    mag, _ = stft_transform(y, fwd, n_fft, hop)
    mel = torch.matmul(torch.from_numpy(basis), mag)
    return torch.log(torch.clamp(mel, min=1e-5)), mag


# --------------------------------------------------------------------------
# clip preprocessing of the loader  (MultiDataset.build_tensor,
# src/data/vid_aud_grid.py:94-121; LRS src/data/vid_aud_lrs2.py:87-120).
# The pixel arithmetic is Pillow's (third-party, un-vendored; 12.2.0 in this image:
# src/libImaging/Resample.c bilinear 8bpc, Convert.c rgb2l), restated in
# integers; pinned bit-exactly to the reference's own build_tensor through
# tests/golden/golden_preproc.npz and, where PIL/torchvision are installed, live.
# --------------------------------------------------------------------------
PIL_PRECISION_BITS = 22


def pil_bilinear_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter.
    -> (first input index [out], tap count [out], int64 coefficients [out, ksize])."""
    scale = in_size / out_size
    fscale = scale if scale > 1.0 else 1.0
    support = fscale                       # bilinear support 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    first = np.zeros(out_size, np.int64); count = np.zeros(out_size, np.int64)
    coef = np.zeros((out_size, ksize), np.int64)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = 0 if lo < 0 else lo
        hi = int(center + support + 0.5)
        hi = in_size if hi > in_size else hi
        weights = []
        for x in range(lo, hi):
            a = (x - center + 0.5) * (1.0 / fscale)
            a = -a if a < 0.0 else a
            weights.append(1.0 - a if a < 1.0 else 0.0)
        s = 0.0
        for w in weights:
            s += w
        for i, w in enumerate(weights):
            w = w / s if s != 0.0 else w
            coef[o, i] = int(0.5 + w * (1 << PIL_PRECISION_BITS)) if w >= 0 else int(-0.5 + w * (1 << PIL_PRECISION_BITS))
        first[o], count[o] = lo, hi - lo
    return first, count, coef


def _pil_resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8bpc pass along axis 0 of a (n, ..., 3) uint8 array: 2^21 + sum(p * k) >> 22, clipped to a byte."""
    first, count, coef = pil_bilinear_tables(img.shape[0], out_size)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    for o in range(out_size):
        acc = np.full(img.shape[1:], 1 << (PIL_PRECISION_BITS - 1), np.int64)
        for j in range(int(count[o])):
            acc += src[first[o] + j] * coef[o, j]
        out[o] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    return out


def preprocess_clip(frames: np.ndarray, boxes: np.ndarray, max_t: int, flip: bool = False,
                    erase: Optional[Tuple[int, int]] = None, out_size: int = 112) -> torch.Tensor:
    """build_tensor (vid_aud_grid.py:94-121): frames uint8 (n,H,W,3), boxes (n,4) or (4,) = left, upper, right, lower
    -> float32 (1, max_t, out, out).  Crop (zero outside the frame) -> Resize (horizontal pass, then vertical) ->
    hflip -> luma -> /255 -> Normalize(0.4136, 0.17); frames n..max_t-1 stay zero; erase = (x_s, y_s) of :116-117."""
    boxes = np.broadcast_to(np.asarray(boxes).reshape(-1, 4), (len(frames), 4))
    vol = torch.zeros(max_t, 1, out_size, out_size)
    for i, (frame, (l, u, r, b)) in enumerate(zip(frames, boxes)):
        H, W, _ = frame.shape
        crop = np.zeros((b - u, r - l, 3), np.uint8)
        y0, y1, x0, x1 = max(u, 0), min(b, H), max(l, 0), min(r, W)
        if y1 > y0 and x1 > x0:
            crop[y0 - u:y1 - u, x0 - l:x1 - l] = frame[y0:y1, x0:x1]
        horiz = _pil_resample_axis0(crop.transpose(1, 0, 2), out_size).transpose(1, 0, 2)
        img = _pil_resample_axis0(horiz, out_size).astype(np.int64)
        if flip:
            img = img[:, ::-1]
        luma = (img[..., 0] * 19595 + img[..., 1] * 38470 + img[..., 2] * 7471 + 0x8000) >> 16
        t = torch.from_numpy(luma.astype(np.uint8)).to(torch.float32).div(255)           # ToTensor
        vol[i, 0] = (t - torch.tensor(0.4136)) / torch.tensor(0.1700)                    # Normalize
    if erase is not None:
        xs, ys = erase
        vol[:, :, max(0, ys):min(out_size, ys + 56), max(0, xs):min(out_size, xs + 56)] = 0.0
    return vol.transpose(1, 0)


# --------------------------------------------------------------------------
# one G+D training step  (train.py:166-237; LRS variant train_LRS.py:179-243)
# --------------------------------------------------------------------------
MODULES = ("v_front", "gen", "post", "dis1", "dis2", "dis3", "s_dis")


def bilinear_half(mel: torch.Tensor, scale: float) -> torch.Tensor:
    """F.interpolate(mel, scale_factor=scale, mode='bilinear'), train.py:170-171."""
    return F.interpolate(mel, scale_factor=scale, mode="bilinear")


def _params(sd: SD) -> List[torch.Tensor]:
    return [t for t in sd.values() if t.is_floating_point() and t.requires_grad]


def train_step_with_adam(sds: Dict[str, SD], batch: Dict[str, torch.Tensor], noise: torch.Tensor,
                         g_opt: torch.optim.Optimizer, d_opt: torch.optim.Optimizer,
                         lrs: bool = False) -> Dict[str, object]:
    """Exact schedule of train.py:166-237 with dropout disabled and generator noise injected.
    batch: mel (B,1,80,4T), spec (B,1,321,4T), vid (B,1,T,112,112), vid_len (B,) ints.
    Mutates ``sds`` (weights, BN running stats) through the two optimizers."""
    vf, gen, post = sds["v_front"], sds["gen"], sds["post"]
    d1, d2, d3, sdis = sds["dis1"], sds["dis2"], sds["dis3"], sds["s_dis"]
    mel, spec, vid, vid_len = batch["mel"], batch["spec"], batch["vid"], batch["vid_len"]
    for m in ("v_front", "gen", "post"):  # train.py:168
        for p in _params(sds[m]):
            p.grad = None
    mel1 = bilinear_half(mel, 0.25)
    mel2 = bilinear_half(mel, 0.5)
    phon, sent = visual_front(vf, vid, True)
    g1, g2, g3 = decoder(gen, sent, phon, vid_len, noise, True)
    T = phon.size(1)
    mel_r = mel.detach().clone().requires_grad_(True)
    mel1_r = mel1.detach().clone().requires_grad_(True)
    mel2_r = mel2.detach().clone().requires_grad_(True)
    sdet = sent.detach()
    ur1, cr1 = discriminator(d1, mel1_r, sdet, T)
    ur2, cr2 = discriminator(d2, mel2_r, sdet, T)
    ur3, cr3 = discriminator(d3, mel_r, sdet, T)
    sync_loss = sync_discriminator(sdis, phon, mel_r, False, True).mean()
    gr1 = torch.autograd.grad(ur1.sum(), mel1_r, create_graph=True)[0]
    gr2 = torch.autograd.grad(ur2.sum(), mel2_r, create_graph=True)[0]
    gr3 = torch.autograd.grad(ur3.sum(), mel_r, create_graph=True)[0]
    gp = [(g.reshape(g.size(0), -1).norm(2, dim=1) ** 2).mean() for g in (gr1, gr2, gr3)]
    uf1, cf1 = discriminator(d1, g1.detach(), sdet, T)
    uf2, cf2 = discriminator(d2, g2.detach(), sdet, T)
    uf3, cf3 = discriminator(d3, g3.detach(), sdet, T)
    real_loss = (1 / 3) * sum(gan_loss(x, True) for x in (ur1, ur2, ur3, cr1, cr2, cr3)) + (1 / 3) * sum(gp)
    fake_loss = (1 / 3) * sum(gan_loss(x, False) for x in (uf1, uf2, uf3, cf1, cf2, cf3))
    sync_w = 0.5 if lrs else 1.0  # train_LRS.py:218
    dis_loss = real_loss + fake_loss + sync_w * sync_loss
    d_opt.zero_grad()
    dis_loss.backward(retain_graph=True)
    d_grad_norms = {m: {k: float(v.grad.norm()) for k, v in sds[m].items()
                        if v.is_floating_point() and v.requires_grad and v.grad is not None}
                    for m in ("dis1", "dis2", "dis3", "s_dis")}
    vf_grad_after_d = {k: v.grad.detach().clone() for k, v in vf.items()
                       if v.is_floating_point() and v.requires_grad and v.grad is not None}
    d_opt.step()
    # ---- G phase ----
    gs = postnet(post, g3, True)
    ug1, cg1 = discriminator(d1, g1, sdet, T)
    ug2, cg2 = discriminator(d2, g2, sdet, T)
    ug3, cg3 = discriminator(d3, g3, sdet, T)
    g_sync = sync_discriminator(sdis, phon.detach(), g3, True, True).mean()
    g_adv = (1 / 3) * sum(gan_loss(x, True) for x in (ug1, ug2, ug3, cg1, cg2, cg3))
    if lrs:  # train_LRS.py:233-237 -- L1 on raw mels, sync added outside g_loss
        recon = (F.l1_loss(g1, mel1) + F.l1_loss(g2, mel2) + F.l1_loss(g3, mel)) / 3.0 + F.l1_loss(gs, spec)
    else:    # train.py:226-229 -- L1 on de-normalised mels
        recon = (F.l1_loss(denormalize(g1), denormalize(mel1)) + F.l1_loss(denormalize(g2), denormalize(mel2))
                 + F.l1_loss(denormalize(g3), denormalize(mel))) / 3.0 + F.l1_loss(gs, spec)
    g_loss = g_adv + g_sync
    gen_loss = g_loss + 50.0 * recon
    for m in ("dis1", "dis2", "dis3", "s_dis", "gen", "post"):  # train.py:235 (v_front NOT zeroed)
        for p in _params(sds[m]):
            p.grad = None
    gen_loss.backward()
    g_grad_norms = {m: {k: float(v.grad.norm()) for k, v in sds[m].items()
                        if v.is_floating_point() and v.requires_grad and v.grad is not None}
                    for m in ("v_front", "gen", "post")}
    g_opt.step()
    return dict(dis_loss=dis_loss.detach(), sync_loss=sync_loss.detach(), real_loss=real_loss.detach(),
                fake_loss=fake_loss.detach(), grad_pen=torch.stack([g.detach() for g in gp]),
                gen_loss=gen_loss.detach(), g_adv=g_adv.detach(), g_sync=g_sync.detach(), recon=recon.detach(),
                g1=g1.detach(), g2=g2.detach(), g3=g3.detach(), gs=gs.detach(),
                r1_grads=[gr1.detach(), gr2.detach(), gr3.detach()],
                d_grad_norms=d_grad_norms, g_grad_norms=g_grad_norms, vf_grad_after_d=vf_grad_after_d)


# --------------------------------------------------------------------------
# deterministic weights shared by the golden generator, the oracle tests and the CUDA parity tests
# --------------------------------------------------------------------------
def det_tensor(name: str, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
    """Name-keyed deterministic parameter/buffer values (independent of constructor RNG order)."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    leaf = name.rsplit(".", 1)[-1]
    if dtype in (torch.int64, torch.int32):
        return torch.zeros(shape, dtype=dtype)
    r = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if leaf == "running_var":
        return 1.0 + 0.2 * r.abs()
    if leaf == "running_mean":
        return 0.1 * r
    if len(shape) == 1:
        if leaf == "bias" or "bias_" in leaf:
            return 0.05 * r
        if "relu" in name or name.endswith("frontend.2.weight") or name.endswith("frontend.5.weight"):
            return 0.25 + 0.05 * r  # PReLU slopes
        return 1.0 + 0.1 * r  # BN gamma
    fan_in = 1
    for s in shape[1:]:
        fan_in *= int(s)
    return r * (1.0 / math.sqrt(fan_in))


def fill_deterministic(sd: SD, prefix: str) -> SD:
    """Return a new dict with every entry of ``sd`` replaced by det_tensor(prefix+key)."""
    return {k: det_tensor(prefix + "." + k, v.shape, v.dtype) for k, v in sd.items()}
