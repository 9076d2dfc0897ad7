#!/usr/bin/env python
"""Time the clip preprocessing kernel at config 2 (B = 32, T = 75 frames of 288x360x3 uint8 -> 112x112 fp32) and at the
LRS shape (16 x 250 frames of 160x160, 80x80 boxes).  Algorithmic bytes = crop window read once + fp32 output written
once.  Prints one JSON line per shape."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from vcagan_b200 import preprocess as P
from vcagan_b200.preprocess import preprocess_clips


def ev_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6531.9)
    for name, (B, T, H, W), crop in (("GRID B=32 T=75 288x360", (32, 75, 288, 360), (59, 95, 195, 231)),
                                     ("LRS B=16 T=250 160x160", (16, 250, 160, 160), None)):
        frames = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8, device="cuda")
        if crop is None:
            c = np.random.default_rng(0).integers(40, 120, (B, T, 2))
            crop = np.concatenate([c - 40, c + 40], -1)
            cw = 80
        else:
            cw = crop[2] - crop[0]
        ms = ev_time(lambda: preprocess_clips(frames, crop=crop))
        meta = np.zeros((B, T, 10), np.int32)
        meta[..., 0:4] = np.broadcast_to(np.asarray(crop).reshape((1, 1, 4) if np.ndim(crop) == 1 else (B, T, 4)), (B, T, 4))
        meta[..., 9] = 1
        meta_d = torch.from_numpy(meta).cuda()
        ms_k = ev_time(lambda: P._launch(frames, meta_d, cw, cw))
        alg = B * T * (cw * cw * 3 + 112 * 112 * 4)
        print(json.dumps({"shape": name, "ms_api_incl_host_descriptors": ms, "ms_kernel": ms_k,
                          "frames_per_s": B * T / ms * 1e3, "algorithmic_bytes": alg,
                          "roofline": {"bound": "hbm", "achieved": alg / ms_k / 1e6, "peak": peak, "unit": "GB/s",
                                       "frac": alg / ms_k / 1e6 / peak}}))


if __name__ == "__main__":
    main()
