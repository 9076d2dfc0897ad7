#!/usr/bin/env python
"""Two Visual_front instances with identical weights on the same input (bf16, train mode, no dropout): which op is the first
whose output differs between the instances?  (tools/g_noise_probe.py showed 2e-3 between instances, 0 between two calls of
one instance.)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "visual-context-attentional-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import make_state, GOLD
from oracle import vca_oracle as O
import vcagan_b200 as V
from vcagan_b200 import models as M, ops

V.set_precision("bf16")
spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
mods = []
for i in range(2):
    m = M.Visual_front(1); m.load_state_dict(make_state(spec, "v_front")); m.cuda().train()
    m.dropout.p = 0.0; m.sentence_encoder.dropout = 0.0
    mods.append(m)
g = torch.Generator().manual_seed(5)
vid = torch.randn(2, 1, 20, 112, 112, generator=g).cuda()
logs = []
names = ["conv", "bn_act", "stem_conv", "bn_prelu_maxpool", "maxpool3x3s2", "spatial_mean", "gru_layer", "linear"]
orig = {n: getattr(ops, n) for n in names if hasattr(ops, n)}
cur = []


def wrap(n, f):
    def g2(*a, **k):
        y = f(*a, **k)
        t = y[0] if isinstance(y, tuple) else y
        cur.append((n, t.detach().float().clone()))
        return y
    return g2


for n, f in orig.items():
    setattr(ops, n, wrap(n, f))
outs = []
for rep in range(2):
    for m in mods:
        cur = []
        with torch.no_grad():
            y = m(vid)
        torch.cuda.synchronize()
        logs.append(cur)
a, b, a2 = logs[0], logs[1], logs[2]
print("ops per forward:", len(a))
first = True
for i, ((n, x), (_, y), (_, z)) in enumerate(zip(a, b, a2)):
    d = float((x - y).norm() / (x.norm() + 1e-30)); d2 = float((x - z).norm() / (x.norm() + 1e-30))
    if d > 0 or d2 > 0 or i < 3:
        print(i, n, tuple(x.shape), "inst0 vs inst1:", d, " inst0 call1 vs call2:", d2)
        if d > 0 and first:
            first = False
