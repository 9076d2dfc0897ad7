"""Clip preprocessing on the B200: the per-frame torchvision / PIL pipeline of
`MultiDataset.build_tensor` (src/data/vid_aud_grid.py:94-121; LRS: src/data/vid_aud_lrs2.py:87-120) for a whole batch
of raw uint8 frames in one kernel launch, bit-exact with the reference's CPU result.

    vid = preprocess_clips(frames_u8)                                   # GRID: fixed crop [59, 95, 195, 231]
    vid = preprocess_clips(frames_u8, crop=boxes, n_frames=lengths)     # LRS: per-frame 80x80 boxes, ragged clips
    vid = preprocess_clips(frames_u8, flip=flags, erase=starts)         # training augmentations (:96-97, :116-118)

The host computes what the reference computes once per call on the host as well -- PIL's resampling coefficient
tables (double precision, then 22-bit fixed point) -- and a 10-int descriptor per frame; the pixels only ever exist on
the device.  There is no CPU fallback."""
import functools
import math
from typing import Optional, Sequence, Union

import numpy as np
import torch

from ._lib import lib

GRID_CROP = (59, 95, 195, 231)          # vid_aud_grid.py:99 (left, upper, right, lower)
MEAN, STD = 0.4136, 0.1700              # :109
ERASE = 56                              # :117-118
_PRECISION_BITS = 32 - 8 - 2            # Pillow Resample.c, 8 bits per channel


@functools.lru_cache(maxsize=32)
def resize_coeffs(in_size: int, out_size: int):
    """PIL `Image.resize(..., BILINEAR)` tables for one axis: (k int32 [out][ksize], bounds int32 [out][2] = first input
    index, tap count).  Follows Pillow's precompute_coeffs + normalize_coeffs_8bpc operation by operation in double
    precision (triangle filter, support 1 scaled by the shrink factor, window clipped to the image, weights
    normalised to sum 1, rounded to 22-bit fixed point)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    k = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) * inv)) for x in range(xmax)]
        total = 0.0
        for v in w:
            total += v
        for x, v in enumerate(w):
            if total != 0.0:
                v = v / total
            k[xx, x] = int(v * (1 << _PRECISION_BITS) + (0.5 if v >= 0 else -0.5))
        bounds[xx] = (xmin, xmax)
    return k, bounds


_dev_tables = {}


def _tables(in_w, in_h, out_w, out_h, device):
    key = (in_w, in_h, out_w, out_h, str(device))
    t = _dev_tables.get(key)
    if t is None:
        kx, bx = resize_coeffs(in_w, out_w)
        ky, by = resize_coeffs(in_h, out_h)
        t = tuple(torch.from_numpy(a).to(device) for a in (kx, bx, ky, by)) + (kx.shape[1], ky.shape[1])
        _dev_tables[key] = t
    return t


def preprocess_clips(frames: torch.Tensor, crop: Union[Sequence[int], torch.Tensor, np.ndarray] = GRID_CROP,
                     flip: Optional[Sequence[bool]] = None, erase: Optional[Sequence[Sequence[int]]] = None,
                     n_frames: Optional[Sequence[int]] = None, out_size: int = 112, mean: float = MEAN,
                     std: float = STD) -> torch.Tensor:
    """frames: uint8 CUDA tensor (B,T,H,W,3) of decoded RGB frames (torchvision.io.read_video layout, vid_aud_grid.py:127)
    -> (B,1,T,out,out) fp32, the `vid` tensor of the training batch.

    crop     one (left, upper, right, lower) box for every frame, or an integer array (B,T,4); all boxes must have the
             same width and height; parts outside the frame read as black, as PIL's crop does.
    flip     per clip: mirror horizontally (StatefulRandomHorizontalFlip draws once per clip, transforms.py:4-12).
    erase    per clip: (x_s, y_s) as drawn by `random.randint(-10, 66)` (vid_aud_grid.py:116); the 56x56 box clipped to
             the image is zeroed in every frame.  None / a negative-size box: no erasing.
    n_frames per clip: frames at t >= n_frames[b] are zero (temporalVolume is zero-initialised, :112)."""
    if not frames.is_cuda:
        raise RuntimeError("preprocess_clips needs CUDA tensors: there is no CPU fallback")
    if frames.dtype != torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3:
        raise ValueError("frames must be uint8 (B,T,H,W,3)")
    frames = frames.contiguous()
    B, T, H, W, _ = frames.shape
    boxes = np.asarray(crop.cpu() if torch.is_tensor(crop) else crop, dtype=np.int64)
    boxes = np.broadcast_to(boxes.reshape((1, 1, 4) if boxes.ndim == 1 else (B, T, 4)), (B, T, 4))
    cw, ch = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
    if (cw != cw.flat[0]).any() or (ch != ch.flat[0]).any() or cw.flat[0] <= 0 or ch.flat[0] <= 0:
        raise ValueError("all crop boxes must have the same positive width and height")
    cw, ch = int(cw.flat[0]), int(ch.flat[0])
    meta = np.zeros((B, T, 10), dtype=np.int32)
    meta[..., 0:4] = boxes
    if flip is not None:
        meta[..., 4] = np.asarray(flip, dtype=bool).reshape(B, 1)
    if erase is not None:
        e = np.asarray(erase, dtype=np.int64).reshape(B, 2)
        meta[..., 5] = np.maximum(0, e[:, 0])[:, None]
        meta[..., 6] = np.maximum(0, e[:, 1])[:, None]
        meta[..., 7] = np.minimum(out_size, e[:, 0] + ERASE)[:, None]
        meta[..., 8] = np.minimum(out_size, e[:, 1] + ERASE)[:, None]
    lens = np.full(B, T) if n_frames is None else np.asarray(
        n_frames.cpu() if torch.is_tensor(n_frames) else n_frames, dtype=np.int64).reshape(B)
    meta[..., 9] = np.arange(T)[None, :] < lens[:, None]
    meta_d = torch.from_numpy(meta).to(frames.device, non_blocking=True)
    return _launch(frames, meta_d, cw, ch, out_size, mean, std)


def _launch(frames, meta_d, cw, ch, out_size=112, mean=MEAN, std=STD):
    """The device part: frames (B,T,H,W,3) uint8 + descriptors int32 (B,T,10) -> (B,1,T,out,out) fp32."""
    B, T, H, W, _ = frames.shape
    kx, bx, ky, by, ksx, ksy = _tables(cw, ch, out_size, out_size, frames.device)
    out = torch.empty((B, 1, T, out_size, out_size), dtype=torch.float32, device=frames.device)
    lib().call("vca_clip_preprocess", frames, B * T, H, W, meta_d, kx, bx, ky, by, ksx, ksy, cw, ch, out_size, out_size,
               float(mean), float(std), out)
    return out
