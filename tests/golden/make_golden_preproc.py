"""Generate tests/golden/golden_preproc.npz: clip preprocessing computed by the
UNMODIFIED reference `MultiDataset.build_tensor` of src/data/vid_aud_grid.py:94-121 and src/data/vid_aud_lrs2.py:87-120
(torchvision transforms over PIL images), with and without the training augmentations.

Run in the build container only (needs /root/reference, torchvision and PIL):
    python tests/golden/make_golden_preproc.py
Shims: librosa / matplotlib stubs as in make_golden_tail.py (imported by the dataset modules, unused by build_tensor).
The dataset objects are created without __init__; the Python `random` module is seeded before each augmented call and
the draws (flip decision, erase start, LRS crop jitter) are reproduced here in the reference's order and stored next to
the outputs.  Input frames come from conftest.synthetic_frames (seeded), so only outputs are committed.
"""
import os, random, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("VCA_REFERENCE", "/root/reference")

from conftest import synthetic_frames, PKG  # noqa: E402
# conftest puts this repo's drop-in `src` package on sys.path; a regular package beats the reference's namespace
# package `src` whatever the order, so take it off again before importing the reference
sys.path[:] = [p for p in sys.path if p != PKG]
sys.path.insert(0, REF)

lib = types.ModuleType("librosa"); util = types.ModuleType("librosa.util"); filt = types.ModuleType("librosa.filters")
util.pad_center = lambda data, size, **k: data
util.tiny = lambda x: np.finfo(np.float32).tiny
util.normalize = lambda x, norm=None, **k: x
filt.mel = lambda *a, **k: np.zeros((80, 321), np.float32)
lib.util = util; lib.filters = filt
mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
mpl.use = lambda *a, **k: None
sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt, "matplotlib": mpl,
                    "matplotlib.pyplot": plt})

import src.data.vid_aud_grid as grid  # noqa: E402
import src.data.vid_aud_lrs2 as lrs2  # noqa: E402


def main():
    out = {}
    # ---- GRID: fixed crop [59, 95, 195, 231] of 288x360 frames, 3 frames in a 5-frame volume ----
    frames = synthetic_frames(11, 3, 288, 360)
    tchw = torch.from_numpy(frames).permute(0, 3, 1, 2)           # vid_aud_grid.py:145 (T C H W)
    ds = object.__new__(grid.MultiDataset)
    ds.max_v_timesteps = 5
    ds.augmentations = False
    out["grid_plain"] = ds.build_tensor(tchw)
    ds.augmentations = True
    for seed in (3, 4, 10):                                        # seeds covering flip / no flip and clipped boxes
        random.seed(seed)
        vol = ds.build_tensor(tchw)
        random.seed(seed)
        flip = random.random() < 0.5                               # StatefulRandomHorizontalFlip.__init__, transforms.py:7
        xs, ys = [random.randint(-10, 66) for _ in range(2)]       # vid_aud_grid.py:116
        out[f"grid_aug{seed}"] = vol
        out[f"grid_aug{seed}_draws"] = np.array([int(flip), xs, ys])

    # ---- LRS: per-frame 80x80 boxes around a landmark, some reaching outside the 160x160 frame ----
    frames = synthetic_frames(12, 4, 160, 160)
    tchw = torch.from_numpy(frames).permute(0, 3, 1, 2)
    centres = [80, 80, 30, 50, 150, 140, 75, 0]                    # (x, y) per frame, vid_aud_lrs2.py:93-98
    dl = object.__new__(lrs2.MultiDataset)
    dl.max_v_timesteps = 6
    dl.augmentations = False
    out["lrs_plain"] = dl.build_tensor(tchw, centres)
    out["lrs_centres"] = np.array(centres)
    dl.augmentations = True
    for seed in (1, 2):
        random.seed(seed)
        vol = dl.build_tensor(tchw, centres)
        random.seed(seed)
        s = random.randint(-5, 5)                                  # vid_aud_lrs2.py:89
        flip = random.random() < 0.5                               # :102
        out[f"lrs_aug{seed}"] = vol
        out[f"lrs_aug{seed}_draws"] = np.array([s, int(flip)])

    path = os.path.join(HERE, "golden_preproc.npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in out.items()})
    print({k: tuple(np.asarray(v).shape) for k, v in out.items()})
    print("wrote", len(out), "arrays;", os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
