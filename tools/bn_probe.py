#!/usr/bin/env python
"""Bandwidth probe of the BatchNorm / column-sum kernels at the step's main shapes (bf16, channels-last [R, C]).
Prints achieved algorithmic GB/s per kernel next to a device copy of the same tensor.   python tools/bn_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
from vcagan_b200._lib import lib

L = lib()
dev = torch.device("cuda")
SHAPES = [(2400 * 56 * 56, 64, 2, 0), (2400 * 28 * 28, 64, 2, 1), (2400 * 14 * 14, 128, 2, 1), (2400 * 7 * 7, 256, 2, 0),
          (32 * 20 * 75, 512, 1, 0), (32 * 40 * 150, 64, 1, 0), (32 * 80 * 300, 32, 1, 0)]
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for R, C, act, has_res in SHAPES:
    x = torch.randn(R, C, device=dev).bfloat16()
    dy = torch.randn(R, C, device=dev).bfloat16()
    res = torch.randn(R, C, device=dev).bfloat16() if has_res else None
    y, dx = torch.empty_like(x), torch.empty_like(x)
    dres = torch.empty_like(x) if has_res else None
    f32 = lambda n=C: torch.empty(n, dtype=torch.float32, device=dev)
    mean, invstd, gamma, beta, pw, rm, rv = f32(), f32(), torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.full((C,), 0.25, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)
    sums = torch.empty(3 * C, dtype=torch.float64, device=dev)
    dg, db, dp = f32(), f32(), f32()
    nb = x.numel() * 2 / 1e6   # MB per tensor pass
    t_copy = timed(lambda: y.copy_(x))
    t_st = timed(lambda: L.call("vca_bn_stats", 1, x, R, C, 1e-5, 0.1, sums, 0, mean, invstd, rm, rv))
    t_fw = timed(lambda: L.call("vca_bn_act_fwd", 1, x, res, y, R, C, mean, invstd, gamma, beta, act, 0.2, pw))
    bw = lambda: L.call("vca_bn_act_bwd", 1, dy, x, res, dx, dres, R, C, mean, invstd, gamma, beta, act, 0.2, pw, 1, sums, dg, db, dp, 0)
    L.cdll.vca_set_option(b"bn_vec", 8)
    t_bw8 = timed(bw)
    L.cdll.vca_set_option(b"bn_vec", 4)
    t_bw = timed(bw)
    t_cs = timed(lambda: L.call("vca_colsum", 1, x, R, C, sums, dg, 0))
    nres = 1 if has_res else 0
    print(f"R={R:>9} C={C:>3} act={act} res={has_res} ({nb:7.1f} MB/pass)  copy {2 * nb / t_copy:7.0f} GB/s | stats {t_st:6.3f} ms {nb / t_st:6.0f} GB/s | "
          f"fwd {t_fw:6.3f} ms {(2 + nres) * nb / t_fw:6.0f} GB/s | bwd(v4) {t_bw:6.3f} ms {(5 + 3 * nres) * nb / t_bw:6.0f} GB/s (v8) {t_bw8:6.3f} ms | "
          f"colsum {t_cs:6.3f} ms {nb / t_cs:6.0f} GB/s", flush=True)
