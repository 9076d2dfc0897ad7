#!/usr/bin/env python
"""Per-step wall time of the pipelined end-to-end loop (Trainer.stage_inputs / replay_prefetched).   python tools/e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200.trainer import Trainer

B, T = 32, 75
dev = torch.device("cuda")
torch.manual_seed(0); V.manual_seed(0)
tr = Trainer(precision="bf16", dropout=True, device=dev)
g = torch.Generator().manual_seed(3)
vid_h = torch.randn(B, 1, T, 112, 112, generator=g).pin_memory()
mel_h = (torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1).pin_memory()
spec_h = torch.rand(B, 1, 321, 4 * T, generator=g).pin_memory()
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
tr.capture(vid_h.to(dev), mel_h.to(dev), spec_h.to(dev), lens)
for mode in ("prefetched", "blocking"):
    torch.cuda.synchronize()
    ts = []
    t0 = time.perf_counter()
    if mode == "prefetched":
        tr.stage_inputs(vid_h, mel_h, spec_h)
    for i in range(24):
        if mode == "prefetched":
            out = tr.replay_prefetched((vid_h, mel_h, spec_h) if i + 1 < 24 else None)
        else:
            out = tr.replay(vid_h, mel_h, spec_h)
        _ = torch.stack([out["gen_loss"], out["dis_loss"]]).cpu()
        t1 = time.perf_counter(); ts.append((t1 - t0) * 1e3); t0 = t1
    print(mode, " ".join(f"{t:.1f}" for t in ts))
