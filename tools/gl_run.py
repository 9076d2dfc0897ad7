#!/usr/bin/env python
"""Three Griffin-Lim iterations at the config-5 size (64 clips x 300 frames): the short command the ncu captures of the
gl_* kernels under profiles/ were taken with (`ncu --set full -k regex:gl_ python tools/gl_run.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
from vcagan_b200 import audio

spec = torch.rand(64, 321, 300, device="cuda")
w = audio.griffin_lim(spec, None, 3)
torch.cuda.synchronize()
print(tuple(w.shape))
