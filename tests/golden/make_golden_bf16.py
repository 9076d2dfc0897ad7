"""Gradient yard-stick of the bf16 bound: per-parameter gradients of ONE reference training step in fp64 (the truth),
fp32 and under torch.autocast(bfloat16), from the UNMODIFIED reference modules driven by the stock step body.

Run in the build container only (needs /root/reference via baseline/_ref):
    python baseline/install_reference.py && python tests/golden/make_golden_bf16.py

Same weights / inputs / injected noise as tests/golden/make_golden.py (B = 2, T = 20, lens 20 / 13, dropout off).
100 M gradient entries per run do not fit a fixture, so every parameter is SAMPLED at up to 512 fixed positions
(`sample_index`); tests/test_gpu_step.py::test_step_bf16_gradient_bound samples the CUDA trainer's gradients at the same
positions and requires, per parameter, an error against the fp64 truth of at most 2 x what the reference itself incurs
under bf16 autocast (plus post-Adam update-sign agreement).  Writes golden_bf16_grads.npz + golden_bf16_names.json.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from baseline import stock_step as S  # noqa: E402
from conftest import sample_index  # noqa: E402
from oracle import vca_oracle as O  # noqa: E402

NS = 512
LENS = [20, 13]
D_MODS, G_MODS = ("dis1", "dis2", "dis3", "s_dis"), ("v_front", "gen", "post")


def gen_inputs(B=2, T=20):
    g = torch.Generator().manual_seed(1234)          # tests/golden/make_golden.py::gen_inputs
    vid = torch.randn(B, 1, T, 112, 112, generator=g)
    mel = torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1
    spec = torch.rand(B, 1, 321, 4 * T, generator=g)
    noise = torch.randn(B, 128, 20, T, generator=g)
    return vid, mel, spec, noise


def run(ns, mode):
    mods = S.build_modules(ns)
    for k, m in mods.items():
        m.load_state_dict(O.fill_deterministic(m.state_dict(), k))
        m.train()
    mods["v_front"].dropout.p = 0.0
    mods["v_front"].sentence_encoder.dropout = 0.0
    vid, mel, spec, noise = gen_inputs()
    dt = torch.float64 if mode == "fp64" else torch.float32
    for m in mods.values():
        m.to(dt)
    vid, mel, spec, noise = (t.to(dt) for t in (vid, mel, spec, noise))
    g_opt, d_opt = S.build_optimizers(mods)
    grads = {}

    def hook(name):
        for k in (D_MODS if name == "d_backward" else G_MODS):
            for n, p in mods[k].named_parameters():
                grads[f"{k}.{n}"] = p.grad.detach().double().reshape(-1)[sample_index(p.numel())].clone()
    orig = torch.randn
    torch.randn = lambda *a, **k: noise.clone()       # generator.py:248 draws the noise with the global torch RNG
    try:
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "autocast")):
            out = S.stock_train_step(mods, g_opt, d_opt, (mel, spec, vid, torch.tensor(LENS)), ns.gan_loss, hook=hook)
    finally:
        torch.randn = orig
    after = {f"{k}.{n}": p.detach().double().reshape(-1)[sample_index(p.numel())].clone()
             for k, m in mods.items() for n, p in m.named_parameters()}
    return grads, after, out


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    ns = S.import_reference(cpu_shim=True)
    res = {m: run(ns, m) for m in ("fp64", "fp32", "autocast")}
    names = sorted(res["fp64"][0])
    out = {}
    for mode, (grads, after, o) in res.items():
        out[f"grad_{mode}"] = np.concatenate([grads[n].numpy() for n in names]).astype(np.float32)
        out[f"after_{mode}"] = np.concatenate([after[n].numpy() for n in names]).astype(np.float64 if mode == "fp64" else np.float32)
        out[f"losses_{mode}"] = np.array([float(o[k]) for k in ("dis_loss", "sync_loss", "gen_loss", "g_sync", "recon")])
    counts = [int(res["fp64"][0][n].numel()) for n in names]
    json.dump(dict(names=names, counts=counts, ns=NS), open(os.path.join(HERE, "golden_bf16_names.json"), "w"))
    np.savez_compressed(os.path.join(HERE, "golden_bf16_grads.npz"), **out)
    # summary for the log
    g64, gac, g32 = (torch.from_numpy(out[f"grad_{m}"]).double() for m in ("fp64", "autocast", "fp32"))
    o, ea, e3 = 0, [], []
    for n, c in zip(names, counts):
        a, b, c3 = g64[o:o + c], gac[o:o + c], g32[o:o + c]; o += c
        den = float(a.norm()) + 1e-300
        ea.append(float((b - a).norm()) / den); e3.append(float((c3 - a).norm()) / den)
    ea, e3 = torch.tensor(ea), torch.tensor(e3)
    print(f"{len(names)} parameters; sampled rel. gradient error vs fp64: autocast-bf16 median {ea.median():.3e} max {ea.max():.3e}; "
          f"fp32 median {e3.median():.3e} max {e3.max():.3e}")
    print("losses", {m: out[f"losses_{m}"].tolist() for m in ("fp64", "fp32", "autocast")})
    print("wrote", os.path.getsize(os.path.join(HERE, "golden_bf16_grads.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
