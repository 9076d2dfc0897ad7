"""Diagnostic (not a test): per-parameter gradient errors of the CUDA modules vs the fp32 oracle and vs an fp64
run of the same oracle (the truth), to separate kernel bugs from fp32 conditioning.  Run on the GPU box:
    python tests/diag_grad_errors.py gen|v_front|s_dis"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import make_state, golden_inputs, rel_l2, GOLD  # noqa
import numpy as np
from oracle import vca_oracle as O
import vcagan_b200 as V

which = sys.argv[1] if len(sys.argv) > 1 else "gen"
spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
golden = dict(np.load(os.path.join(GOLD, "golden_small.npz")))
vid, mel, sp, noise = golden_inputs()
V.set_precision("fp32")


def to64(sd):
    return {k: (v.detach().double().requires_grad_(v.requires_grad) if v.is_floating_point() else v.clone()) for k, v in sd.items()}


def report(mod, sd32, sd64):
    rows = []
    for n, p in mod.named_parameters():
        g32, g64 = sd32[n].grad, sd64[n].grad
        if g64 is None or p.grad is None:
            continue
        rows.append((n, rel_l2(p.grad.cpu(), g64), rel_l2(g32, g64), float(g64.norm())))
    rows.sort(key=lambda r: -r[1])
    print("%-40s %10s %10s %10s" % ("param", "mine-vs-64", "ref32-vs-64", "|g|"))
    for r in rows[:25]:
        print("%-40s %10.2e %10.2e %10.2e" % r)
    print("median mine %.2e  median ref32 %.2e" % (sorted(r[1] for r in rows)[len(rows) // 2], sorted(r[2] for r in rows)[len(rows) // 2]))


if which == "gen":
    sent = torch.from_numpy(golden["eval_sent"]); phon = torch.from_numpy(golden["eval_phon"])
    g = torch.Generator().manual_seed(4)
    outs = {}
    for tag, cast in (("32", lambda t: t), ("64", lambda t: t.double())):
        sd = make_state(spec, "gen", requires_grad=True)
        if tag == "64":
            sd = to64(sd)
        r = O.decoder(sd, cast(sent), cast(phon), [20, 13], cast(noise), True)
        outs[tag] = (sd, r)
    ws = [torch.randn(t.shape, generator=g) for t in outs["32"][1]]
    for tag in ("32", "64"):
        sum((t * w.to(t.dtype)).sum() for t, w in zip(outs[tag][1], ws)).backward()
    m = V.models.Decoder(); m.load_state_dict(make_state(spec, "gen")); m = m.cuda().train(); m.fixed_noise = noise
    o = m(sent.cuda(), phon.cuda(), [20, 13])
    sum((t * w.cuda()).sum() for t, w in zip(o, ws)).backward()
    for a, b32, b64 in zip(o, outs["32"][1], outs["64"][1]):
        print("fwd mine-vs-64 %.2e ref32-vs-64 %.2e" % (rel_l2(a.detach().cpu(), b64), rel_l2(b32, b64)))
    report(m, outs["32"][0], outs["64"][0])
elif which == "v_front":
    g = torch.Generator().manual_seed(2)
    outs = {}
    for tag, cast in (("32", lambda t: t), ("64", lambda t: t.double())):
        sd = make_state(spec, "v_front", requires_grad=True)
        if tag == "64":
            sd = to64(sd)
        outs[tag] = (sd, O.visual_front(sd, cast(vid), True))
    dp, ds = torch.randn(outs["32"][1][0].shape, generator=g), torch.randn(outs["32"][1][1].shape, generator=g)
    for tag in ("32", "64"):
        p_, s_ = outs[tag][1]
        ((p_ * dp.to(p_.dtype)).sum() + (s_ * ds.to(s_.dtype)).sum()).backward()
    m = V.models.Visual_front(); m.load_state_dict(make_state(spec, "v_front")); m = m.cuda().train()
    m.dropout.p = 0.0; m.sentence_encoder.dropout = 0.0
    ph, se = m(vid.cuda())
    ((ph * dp.cuda()).sum() + (se * ds.cuda()).sum()).backward()
    for a, b32, b64 in zip((ph, se), outs["32"][1], outs["64"][1]):
        print("fwd mine-vs-64 %.2e ref32-vs-64 %.2e" % (rel_l2(a.detach().cpu(), b64), rel_l2(b32, b64)))
    report(m, outs["32"][0], outs["64"][0])
