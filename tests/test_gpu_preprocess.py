"""Clip preprocessing kernel (csrc/preprocess.cu) against the reference's
`MultiDataset.build_tensor` vectors (tests/golden/golden_preproc.npz) and the oracle.  Integer pixel arithmetic and
IEEE fp32 division: the bar is BIT-EXACT (torch.equal), also at the full batch size of config 2."""
import numpy as np
import pytest
import torch

from conftest import synthetic_frames
from oracle import vca_oracle as O
from test_oracle_golden import preproc_cases

pytestmark = pytest.mark.gpu


def run(frames, boxes, max_t, flip, erase):
    """One clip through the batch API, padded to max_t frames like the reference's temporalVolume."""
    from vcagan_b200.preprocess import preprocess_clips
    n, H, W, _ = frames.shape
    buf = np.zeros((1, max_t, H, W, 3), np.uint8)
    buf[0, :n] = frames
    buf[0, n:] = 77                                        # garbage behind the clip must not show up
    b = np.zeros((1, max_t, 4), np.int64)
    b[0, :n] = np.broadcast_to(np.asarray(boxes).reshape(-1, 4), (n, 4))
    b[0, n:] = b[0, 0]
    return preprocess_clips(torch.from_numpy(buf).cuda(), crop=b, flip=[flip], erase=None if erase is None else [erase],
                            n_frames=[n])[0].cpu()


def test_matches_reference_build_tensor_bit_exactly(golden_preproc):
    for name, frames, boxes, max_t, flip, erase in preproc_cases(golden_preproc):
        got = run(frames, boxes, max_t, flip, erase)
        assert got.shape == golden_preproc[name].shape
        assert torch.equal(got, torch.from_numpy(golden_preproc[name])), name


def test_default_grid_crop_and_batch_arguments():
    from vcagan_b200.preprocess import preprocess_clips
    frames = synthetic_frames(21, 4, 240, 200).reshape(2, 2, 240, 200, 3)
    out = preprocess_clips(torch.from_numpy(frames).cuda(), flip=[False, True], erase=[(-10, 66), (30, -5)]).cpu()
    assert out.shape == (2, 1, 2, 112, 112) and out.dtype == torch.float32
    for b, (flip, erase) in enumerate(((False, (-10, 66)), (True, (30, -5)))):
        assert torch.equal(out[b], O.preprocess_clip(frames[b], np.array([59, 95, 195, 231]), 2, flip, erase)), b
    with pytest.raises(ValueError):
        preprocess_clips(torch.from_numpy(frames).cuda(), crop=np.array([[[0, 0, 80, 80], [0, 0, 81, 80]]] * 2))
    with pytest.raises(RuntimeError):
        preprocess_clips(torch.from_numpy(frames))


@pytest.mark.parametrize("size,box", [(80, (-40, -40, 40, 40)), (136, (100, 150, 236, 286)), (150, (5, 5, 155, 155)),
                                      (300, (0, 0, 300, 300)), (57, (10, 20, 67, 77)),
                                      (99, (20, 30, 120, 166)), (98, (-5, 10, 395, 110))])
def test_other_crop_sizes_and_out_of_frame_boxes(size, box):
    """Enlarging (80 -> 112, 3 taps), shrinking (up to 300 -> 112, 7 taps, > 48 KB of shared memory), boxes hanging
    over every edge of the frame, non-square boxes (3 x 5 and 9 x 3 taps: the generic kernel)."""
    frames = synthetic_frames(size, 2, 240, 260)
    assert torch.equal(run(frames, np.array(box), 3, False, None), O.preprocess_clip(frames, np.array(box), 3))


def test_full_batch_config2():
    """B = 32 clips x T = 75 frames of 288x360 (GRID), ragged lengths: sampled frames bit-exact against the oracle,
    every frame behind a clip's end is zero, flipped clips are mirror images of the unflipped result."""
    from vcagan_b200.preprocess import preprocess_clips
    B, T, H, W = 32, 75, 288, 360
    base = torch.from_numpy(synthetic_frames(5, 6, H, W)).cuda()
    g = torch.Generator().manual_seed(0)
    pick = torch.randint(0, 6, (B, T), generator=g)
    shift = torch.randint(0, 255, (B, T), generator=g).to(torch.uint8)
    frames = base[pick.cuda()] + shift.cuda()[:, :, None, None, None]            # uint8 wrap-around: distinct frames
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    flip = torch.rand(B, generator=g) < 0.5
    out = preprocess_clips(frames, flip=flip.tolist(), n_frames=lens)
    plain = preprocess_clips(frames, n_frames=lens)
    assert out.shape == (B, 1, T, 112, 112)
    for b in range(B):
        assert float(out[b, :, int(lens[b]):].abs().max()) == 0.0 if lens[b] < T else True
        assert torch.equal(out[b], plain[b].flip(-1) if flip[b] else plain[b])
    fr = frames.cpu().numpy()
    for b, t in ((0, 0), (7, 20), (31, int(lens[31]) - 1), (16, 3)):
        ref = O.preprocess_clip(fr[b, t:t + 1], np.array([59, 95, 195, 231]), 1, bool(flip[b]))
        assert torch.equal(out[b, :, t].cpu(), ref[:, 0]), (b, t)
