#!/usr/bin/env python
"""Stem im2col (2400 frames of 112 x 112 -> [2400, 56, 56, 64] bf16): time per band height ("im2col_rb")."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200._lib import lib
L = lib()
vid = torch.randn(2400, 112, 112, device="cuda")
out = torch.empty(2400, 56, 56, 64, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for rb in (0, 4, 8, 14, 28):
    assert L.cdll.vca_set_option(b"im2col_rb", rb) == 0
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); L.call("vca_stem_im2col", 0, 1, vid, out, 2400, 112, 112); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if ref is None:
        ref = out.clone()
    print(f"im2col_rb={rb}: {sorted(ts)[2]:.3f} ms  ({(vid.numel() * 4 + out.numel() * 2) / sorted(ts)[2] / 1e9:.2f} TB/s)  equal to the untiled kernel: {torch.equal(out, ref)}")
