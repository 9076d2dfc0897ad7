#!/usr/bin/env python
"""BASELINE config 5: batched test-time inference -- generator (+flip TTA) + Postnet + Griffin-Lim (60 iterations) for
64 GRID clips on one B200.  Prints one JSON line with clips/s and the Griffin-Lim HBM roofline fraction.
    python tools/bench_inference.py [--batch 64] [--frames 75] [--iters 60]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200")); sys.path.insert(0, ROOT)
import torch
import vcagan_b200 as V
from vcagan_b200 import models as M, audio, infer


def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64); ap.add_argument("--frames", type=int, default=75)
    ap.add_argument("--iters", type=int, default=60)
    a = ap.parse_args()
    V.set_precision("bf16")
    torch.manual_seed(1)
    dev = torch.device("cuda")
    vf, gen, post = M.Visual_front().to(dev).eval(), M.Decoder().to(dev).eval(), M.Postnet().to(dev).eval()
    B, T = a.batch, a.frames
    vid = torch.randn(B, 1, T, 112, 112, device=dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    spec = torch.rand(B, 321, 4 * T, device=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    ms_gl = ev_time(lambda: audio.griffin_lim(spec, None, a.iters))
    Tp = 4 * T
    alg_bytes = B * (a.iters + 1) * (321 * Tp * 4 + 2 * 160 * (Tp - 1) * 4)   # mag read + signal read + write
    ms_all = ev_time(lambda: infer.synthesize(vf, gen, post, vid, lens, n_iters=a.iters, tta=True))
    ms_fwd = ev_time(lambda: infer.synthesize(vf, gen, post, vid, lens, n_iters=0, tta=True))
    wav = audio.griffin_lim(spec, None, 1)
    ms_de = ev_time(lambda: audio.deemphasize(wav))
    tstft = audio.TacotronSTFT().to(dev)
    mel = torch.rand(B, 1, 80, Tp, device=dev) * 2 - 1
    ms_m2s = ev_time(lambda: tstft.mel_to_spec(mel))
    line = {"metric": "test-time inference clips/s (generator + flip TTA + Postnet + Griffin-Lim)", "value": B / ms_all * 1e3,
            "unit": "clips/s", "batch": B, "frames": T, "gl_iters": a.iters, "ms_total": ms_all, "ms_forward_tta": ms_fwd,
            "ms_griffin_lim": ms_gl,
            "ms_deemphasis_clip": ms_de, "deemphasis_gbs": 2 * 4 * wav.numel() / (ms_de * 1e-3) / 1e9,
            "ms_mel_to_spec": ms_m2s, "dtype": "bf16 network / f32 Griffin-Lim",
            "roofline": {"bound": "hbm", "kernel": "gl_frames_kernel + gl_ola_kernel", "achieved": alg_bytes / (ms_gl * 1e-3) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg_bytes / (ms_gl * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "algorithmic_bytes": alg_bytes}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
