#!/usr/bin/env python
"""Griffin-Lim, 64 clips x 300 frames x 60 iterations: the fused one-kernel-per-iteration path against the two-kernel path
(time with CUDA events, and the difference between their outputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
from vcagan_b200 import audio
from vcagan_b200._lib import lib

g = torch.Generator().manual_seed(5)
B, T = 64, 300
sig = torch.randn(B, 160 * (T - 1), generator=g).cuda() * 0.1
mag, _ = audio.STFT(640, 160, 640).transform(sig)
init = (torch.rand(mag.shape, generator=g) * 2 - 1).mul(3.14159).cuda()
outs = {}
for fpw in (1, 2, 4):
    assert lib().cdll.vca_set_option(b"gl_fpw", fpw) == 0
    for fused in (False, True):
        audio.FUSED_ITERATIONS = fused
        for _ in range(2):
            w = audio.griffin_lim(mag, None, 60, init_angles=init)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); w = audio.griffin_lim(mag, None, 60, init_angles=init); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        outs[(fpw, fused)] = w
        print(f"fpw={fpw} fused={fused}: {sorted(ts)[2]:.3f} ms for 60 iterations, {B} clips", flush=True)
    d = float((outs[(fpw, True)] - outs[(fpw, False)]).norm() / outs[(fpw, False)].norm())
    print(f"fpw={fpw}: fused vs two-kernel output rel diff {d:.3e}")
