#!/usr/bin/env python
"""Griffin-Lim (60 iterations, 64 clips x 300 frames) for a few values of the gl_fpw tuning option."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200 import audio

spec = torch.rand(64, 321, 300, device="cuda")
for fpw in (1, 2, 3, 4, 5, 2):
    assert V.lib().cdll.vca_set_option(b"gl_fpw", fpw) == 0
    audio.griffin_lim(spec, None, 5); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); audio.griffin_lim(spec, None, 60); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("gl_fpw", fpw, "ms", sorted(ts)[1])
