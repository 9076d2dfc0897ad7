#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 conv kernels on the layer shapes of the B=32, T=75 training step (CUDA events on
the launching stream, L2 flushed between launches).  Used for the ncu captures under profiles/.
    python tools/conv_shapes.py [--reps 5] [--only fwd|dgrad|wgrad]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200.ops import _geom, _packed
from vcagan_b200._lib import lib

SHAPES = [  # name, N, H, W, Cin, Cout, (kh,kw), (ph,pw)
    ("gen.decode.0.conv1", 32, 20, 75, 640, 512, (5, 5), (2, 2)),
    ("gen.decode.0.conv2", 32, 20, 75, 512, 512, (5, 5), (2, 2)),
    ("gen.decode.2.conv2", 32, 20, 75, 256, 256, (5, 5), (2, 2)),
    ("gen.g1.1.conv1", 32, 20, 75, 128, 128, (5, 5), (2, 2)),
    ("gen.g2.1.conv1", 32, 40, 150, 64, 64, (5, 5), (2, 2)),
    ("gen.g3.1.conv1", 32, 80, 300, 32, 32, (5, 5), (2, 2)),
    ("resnet.layer1.conv", 2400, 28, 28, 64, 64, (3, 3), (1, 1)),
    ("resnet.layer2.conv", 2400, 14, 14, 128, 128, (3, 3), (1, 1)),
    ("resnet.layer3.conv", 2400, 7, 7, 256, 256, (3, 3), (1, 1)),
    ("resnet.layer4.conv", 2400, 4, 4, 512, 512, (3, 3), (1, 1)),
    ("v_front.stem(5,1)", 32, 75, 3136, 64, 64, (5, 1), (2, 0)),
    ("dis3.cond.1", 32, 5, 18, 1024, 512, (5, 5), (2, 2)),
    ("dis3.uncond.1(p0)", 32, 5, 18, 512, 512, (5, 5), (0, 0)),
    ("gru.proj(linear)", 2400, 1, 1, 1024, 1536, (1, 1), (0, 0)),
    ("gen.g3.pair(5x3)", 32, 80, 150, 64, 64, (5, 3), (2, 1)),
    ("dis.sc(1x1,32-64)", 32, 80, 300, 32, 64, (1, 1), (0, 0)),
    ("dis.sc(1x1,64-128)", 32, 40, 150, 64, 128, (1, 1), (0, 0)),
    ("dis.conv(32-64,5x5)", 32, 40, 150, 32, 64, (5, 5), (2, 2)),
    ("gen.g2.0.conv1(128-64)", 32, 40, 150, 128, 64, (5, 5), (2, 2)),
    ("gen.g3.0.conv1(64-32)", 32, 80, 300, 64, 32, (5, 5), (2, 2)),
    ("gen.attconv2(96-64)", 32, 40, 150, 96, 64, (5, 5), (2, 2)),
    ("gen.g1.x(256-128)", 32, 20, 75, 256, 128, (5, 5), (2, 2)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--filter", default="")
    ap.add_argument("--ws", type=int, default=1, help="weights-stationary kernel: 0 off, 1 auto, 2 force")
    ap.add_argument("--hs", type=int, default=1, help="halo-resident / streamed-weights kernel: 0 off, 1 auto, 2 force")
    ap.add_argument("--splitk", type=int, default=1, help="lend the split-K workspace the library asks for (0 = never split)")
    ap.add_argument("--wgws", type=int, default=1, help="multi-tap wgrad kernel: 0 off, 1 auto (Cin <= 128), 2 any Cin")
    ap.add_argument("--opt", action="append", default=[], help="key=value passed to vca_set_option (repeatable)")
    ap.add_argument("--tag", default="")
    ap.add_argument("--sweep", default="", help="key=v1,v2,...: run everything once per value of this vca_set_option key")
    args = ap.parse_args()
    if args.sweep:
        key, vals = args.sweep.split("=")
        for v in vals.split(","):
            assert lib().cdll.vca_set_option(key.encode(), int(v)) == 0
            args.tag = f"{key}={v} "
            run(args)
        return
    run(args)


def run(args):
    for kv in args.opt:
        k, v = kv.split("=")
        assert lib().cdll.vca_set_option(k.encode(), int(v)) == 0, kv
    assert lib().cdll.vca_set_option(b"ws_mode", args.ws) == 0
    assert lib().cdll.vca_set_option(b"hs_mode", args.hs) == 0
    assert lib().cdll.vca_set_option(b"wgws_mode", args.wgws) == 0
    V.set_precision("bf16")
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, N, H, W, Cin, Cout, k, p in SHAPES:
        if args.filter and args.filter not in name:
            continue
        x = torch.randn(N, H, W, Cin, device=dev).bfloat16()
        w = torch.nn.Parameter(torch.randn(Cout, Cin, *k, device=dev) / (Cin * k[0] * k[1]) ** 0.5)
        g, oshape = _geom(x.shape, w.shape, (1, 1), p)
        y = torch.empty(oshape, dtype=torch.bfloat16, device=dev)
        dy = torch.randn(oshape, device=dev).bfloat16()
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        wf, wd = _packed(w, torch.bfloat16)
        flops = 2 * N * oshape[1] * oshape[2] * Cout * Cin * k[0] * k[1]
        wsf = lib().query("vca_conv_tc_workspace", g, 0) if args.splitk else 0
        wsd = lib().query("vca_conv_tc_workspace", g, 1) if args.splitk else 0
        ws = torch.empty(max(wsf, wsd, 4) // 4, dtype=torch.float32, device=dev)
        calls = dict(fwd=lambda: lib().call("vca_conv_fwd_tc_ws", g, x, wd, None, y, ws if wsf else None, wsf),
                     dgrad=lambda: lib().call("vca_conv_dgrad_tc_ws", g, dy, wf, dx, ws if wsd else None, wsd),
                     wgrad=lambda: lib().call("vca_conv_wgrad_tc", g, dy, x, dw))
        for kind, fn in calls.items():
            if args.only and kind != args.only:
                continue
            fn(); torch.cuda.synchronize()
            ms = []
            for _ in range(args.reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            t = sorted(ms)[len(ms) // 2]
            rows.append(dict(layer=name, kind=kind, ms=round(t, 4), tflops=round(flops / t / 1e9, 1), gflop=round(flops / 1e9, 2)))
            print(f"{args.tag}{name:22s} {kind:6s} {t:8.3f} ms  {flops / t / 1e9:8.1f} TFLOP/s  ({flops / 1e9:.1f} GFLOP)", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "conv_shapes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
