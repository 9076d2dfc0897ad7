#!/usr/bin/env python
"""Probe of the weights-stationary / halo-resident conv kernel: correctness of the shifted-window A descriptors for
both base-offset modes, and speed vs the streaming kernel.   python tools/ws_probe.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import torch.nn.functional as F
import vcagan_b200 as V
from vcagan_b200.ops import _geom, _packed
from vcagan_b200._lib import lib

L = lib()


def opt(k, v):
    assert L.cdll.vca_set_option(k.encode(), int(v)) == 0


V.set_precision("bf16")
dev = torch.device("cuda")
CASES = [(2, 32, 12, 30, 32, (5, 5), (2, 2)), (2, 64, 20, 25, 64, (5, 5), (2, 2)), (3, 64, 28, 28, 64, (3, 3), (1, 1)),
         (2, 64, 9, 50, 64, (5, 1), (2, 0)), (2, 96, 10, 21, 64, (5, 5), (2, 2)), (2, 32, 7, 9, 64, (5, 5), (0, 0))]
for N, Cin, H, W, Cout, k, p in CASES:
    g = torch.Generator().manual_seed(N + Cin + H)
    x = torch.randn(N, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, *k, generator=g) / math.sqrt(Cin * k[0] * k[1])).bfloat16().float()
    b = torch.randn(Cout, generator=g)
    y = F.conv2d(x, w, b, 1, p)
    dy = torch.randn(y.shape, generator=g).bfloat16().float()
    dx = torch.autograd.grad(F.conv2d(x.requires_grad_(True), w, None, 1, p), x, dy)[0]
    xd = x.detach().permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    wp = torch.nn.Parameter(w.cuda())
    geom, oshape = _geom(xd.shape, wp.shape, (1, 1), p)
    wf, wd = _packed(wp, torch.bfloat16)
    dwr = torch.autograd.grad(F.conv2d(x.detach(), w.requires_grad_(True), None, 1, p), w, dy)[0]
    for wm in (0, 1):
        opt("wgws_mode", wm)
        dwd = torch.zeros_like(wp.data)
        L.call("vca_conv_wgrad_tc", geom, dyd, xd, dwd)
        torch.cuda.synchronize()
        print(f"case {(N, Cin, H, W, Cout, k, p)} wgws_mode={wm}: wgrad err {float((dwd.cpu() - dwr).norm() / dwr.norm()):.3e}", flush=True)
    opt("wgws_mode", 1)
    for mode, boff in ((0, 0), (2, 0)):
        opt("ws_mode", mode); opt("ws_base_off", boff)
        yd = torch.zeros(oshape, dtype=torch.bfloat16, device=dev)
        L.call("vca_conv_fwd_tc", geom, xd, wd, b.cuda(), yd)
        dxd = torch.zeros_like(xd)
        L.call("vca_conv_dgrad_tc", geom, dyd, wf, dxd)
        torch.cuda.synchronize()
        ef = float((yd.float().cpu().permute(0, 3, 1, 2) - y).norm() / y.norm())
        ed = float((dxd.float().cpu().permute(0, 3, 1, 2) - dx).norm() / dx.norm())
        print(f"case {(N, Cin, H, W, Cout, k, p)} ws_mode={mode} base_off={boff}: fwd err {ef:.3e} dgrad err {ed:.3e}", flush=True)
opt("ws_mode", 1); opt("ws_base_off", 0)
