#!/usr/bin/env python
"""The kernels whose `ncu --set full` captures are summarised under profiles/ (round 2), one launch each after a warm-up
launch, at the shapes of the B = 32, T = 75 step:
    ncu --set full --clock-control none --import-source on -k regex:'<names>' --launch-skip-before-match 0 \
        -o gpurun_out/r02_kernels python tools/ncu_targets.py
  1 conv_tc_ws_kernel        ResNet layer 1 conv (2400 x 28 x 28, 64 -> 64, 3x3) WITH the BatchNorm statistics epilogue
  2 conv_tc_fwd_kernel       gen.decode.0.conv2 (32 x 20 x 75, 512 -> 512, 5x5): the 100 %-of-cuBLAS-peak streaming kernel
  3 conv_tc_wgrad_ws_kernel  ResNet layer 1 wgrad
  4 att_fwd_tc_kernel        fused visual-context attention, LRS shape (16 x 500 queries x 250 keys)
  5 bmm_tc_kernel            dK = dS^T Q of the same shape
  6 bn_prelu_maxpool_*       fused stem tail forward / backward at 2400 x 56 x 56 x 64
  7 gl_frames_kernel         one Griffin-Lim half-iteration for 64 clips
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200 import audio
from vcagan_b200.ops import _geom, _packed
from vcagan_b200._lib import lib

V.set_precision("bf16")
dev = torch.device("cuda")
L = lib()


def conv_case(N, H, W, Cin, Cout, k, p, stats=False, wgrad=False):
    x = torch.randn(N, H, W, Cin, device=dev).bfloat16()
    w = torch.nn.Parameter(torch.randn(Cout, Cin, *k, device=dev) / (Cin * k[0] * k[1]) ** 0.5)
    g, oshape = _geom(x.shape, w.shape, (1, 1), p)
    y = torch.empty(oshape, dtype=torch.bfloat16, device=dev)
    wf, wd = _packed(w, torch.bfloat16)
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    for _ in range(1 if os.environ.get("VCA_NCU") else 2):
        if wgrad:   # the production path: TMA reduce-add epilogue into the tap-major slab
            L.call("vca_conv_wgrad_tc_tm", g, torch.randn(oshape, device=dev).bfloat16(), x, torch.zeros(k[0] * k[1], Cout, Cin, device=dev))
        elif stats:
            L.call("vca_conv_fwd_tc_stats", g, x, wd, None, y, sums)
        else:
            L.call("vca_conv_fwd_tc_ws", g, x, wd, None, y, None, 0)
    torch.cuda.synchronize()


conv_case(2400, 28, 28, 64, 64, (3, 3), (1, 1), stats=True)
conv_case(32, 20, 75, 512, 512, (5, 5), (2, 2))
conv_case(2400, 28, 28, 64, 64, (3, 3), (1, 1), wgrad=True)
conv_case(32, 40, 150, 64, 64, (5, 5), (2, 2))                 # 64-ch 5x5: two 32-channel CTA columns, five taps per MMA (N = 160)
conv_case(2400, 14, 14, 128, 128, (3, 3), (1, 1), stats=False)  # ResNet layer 2 on the stacked weights-stationary kernel
conv_case(32, 20, 75, 512, 512, (5, 5), (2, 2), wgrad=True)     # streaming wgrad, 256-wide tiles, TMA reduce-add epilogue
B, Tq, S = 16, 500, 250
q, k, v = (torch.randn(B, n, 256, device=dev).bfloat16().requires_grad_(True) for n in (Tq, S, S))
lens = torch.randint(S // 2, S + 1, (B,), device=dev, dtype=torch.int32)
for _ in range(2):
    o = V.ops.attention(q, k, v, lens, 1 / 16)
    o.backward(torch.randn_like(o))
torch.cuda.synchronize()
NF, H, W, C = 2400, 56, 56, 64
x = torch.randn(NF, H, W, C, device=dev).bfloat16().requires_grad_(True)
bn = torch.nn.BatchNorm2d(C).to(dev).train()
pw = torch.nn.Parameter(torch.full((C,), 0.25, device=dev))
for _ in range(2):
    yy = V.ops.bn_prelu_maxpool(x, bn, pw)
    yy.backward(torch.randn_like(yy))
torch.cuda.synchronize()
spec = torch.rand(64, 321, 300, device=dev)
audio.griffin_lim(spec, None, 2)
torch.cuda.synchronize()
print("ncu targets done")
