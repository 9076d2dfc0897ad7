#!/usr/bin/env python
"""Timing of the GRU recurrence back ends at the step's shape (T=75, B=32, H=512, 2 directions).   python tools/gru_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
from vcagan_b200._lib import lib

L = lib()
dev = torch.device("cuda")
T, B, H = 75, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 512
gi = torch.randn(2, T, B, 3 * H, device=dev)
whh = torch.randn(2, 3 * H, H, device=dev) / 16
bhh = torch.randn(2, 3 * H, device=dev) / 16
dout = torch.randn(T, B, 2 * H, device=dev)


def run(cluster):
    L.cdll.vca_set_option(b"gru_cluster", cluster)
    out = torch.empty(T, B, 2 * H, device=dev); gates = torch.empty(2, T, B, 4 * H, device=dev)
    h = torch.zeros(2, 2, B, H, device=dev); bar = torch.empty(1, dtype=torch.int32, device=dev)
    dgi = torch.empty(2, T, B, 3 * H, device=dev); dgh = torch.empty_like(dgi)
    dhc = torch.empty(2, 2, B, H, device=dev); cur = torch.empty(2, B, 3 * H, device=dev); dhz = torch.empty(2, B, H, device=dev)
    f = lambda: L.call("vca_gru_seq_fwd", gi, whh, bhh, h, out, gates, bar, 2, T, B, H)
    bw = lambda: L.call("vca_gru_seq_bwd", dout, whh, gates, out, dgi, dgh, dhc, cur, dhz, bar, 2, T, B, H)
    res = []
    for fn in (f, bw):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
    return res, out, dgi, dgh


import ctypes
info = (ctypes.c_int * 6)()
L.cdll.vca_gru_cluster_query(B, H, info)
print("cluster plan: CTAs/cluster %d, batch rows/cluster %d, clusters needed %d, co-resident fwd %d bwd %d, K slices %d" % tuple(info))
(rf, rb), o1, a1, b1 = run(1)
(cf, cb), o0, a0, b0 = run(0)
L.cdll.vca_set_option(b"gru_cluster", 1)
rel = lambda a, b: float((a - b).norm() / b.norm())
print(f"cluster: fwd {rf:.3f} ms ({rf / T * 1e3:.2f} us/step)  bwd {rb:.3f} ms ({rb / T * 1e3:.2f} us/step)")
print(f"coop   : fwd {cf:.3f} ms ({cf / T * 1e3:.2f} us/step)  bwd {cb:.3f} ms ({cb / T * 1e3:.2f} us/step)")
print(f"cluster vs coop: out {rel(o1, o0):.2e} dgi {rel(a1, a0):.2e} dgh {rel(b1, b0):.2e}")
