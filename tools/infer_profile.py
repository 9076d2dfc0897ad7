#!/usr/bin/env python
"""Per-entry-point time of the eval forward (v_front + gen with flip TTA + post, B = 64, T = 75), every library call
bracketed by CUDA events (serialised), with and without the fused inference epilogues / stem tail.
    python tools/infer_profile.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200 import models as M, infer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 75
V.set_precision("bf16")
dev = torch.device("cuda")
torch.manual_seed(1)
vf, gen, post = M.Visual_front().to(dev).eval(), M.Decoder().to(dev).eval(), M.Postnet().to(dev).eval()
vid = torch.randn(B, 1, T, 112, 112, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
MODES = [("separate passes", False, False), ("epilogue fusion (scale/shift/residual)", True, False), ("BatchNorm folded into weights", False, True)]
if len(sys.argv) > 2:
    MODES = [m for m in MODES if sys.argv[2] in m[0]] or MODES
for label, fused, fold in MODES:
    V.ops.cfg.fuse_eval_epilogue = fused
    V.ops.cfg.fuse_stem_pool = True
    V.ops.cfg.fold_eval_bn = fold
    run = lambda: infer.synthesize(vf, gen, post, vid, lens, n_iters=0, tta=True)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    prof = V.lib().profile_step(run)
    tot = sum(v["ms"] for v in prof.values())
    print(f"\n==== {label}: forward {a.elapsed_time(b):.2f} ms; sum of library calls (serialised) {tot:.2f} ms")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]:
        print(f"  {k:32s} n={v['n']:4d} {v['ms']:8.3f} ms   {v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f} TF/s")
        for g in v["top"][:(24 if k.startswith('vca_conv') or k.startswith('vca_bn_act') else 3)]:
            print(f"       {g[0]:60s} x{g[1]:3d} {g[2]:8.3f} ms {g[3]:8.1f} TF/s")
