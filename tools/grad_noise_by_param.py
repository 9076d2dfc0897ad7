#!/usr/bin/env python
"""Run-to-run difference of every G-group parameter gradient at the bench shape (B = 32, T = 75, bf16): where does the
order-dependent noise enter the backward?  python tools/grad_noise_by_param.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200.trainer import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = 75
g = torch.Generator().manual_seed(3)
vid = torch.randn(B, 1, T, 112, 112, generator=g).cuda()
mel = (torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1).cuda()
spec = torch.rand(B, 1, 321, 4 * T, generator=g).cuda()
lens = torch.full((B,), T, dtype=torch.int32).cuda()
grads, names = [], None
for run in range(2):
    torch.manual_seed(1); V.manual_seed(1)
    tr = Trainer(precision="bf16", dropout=True)
    tr.split_g_backward = True            # the generator sees detached leaves phon_g / sent_g: their gradients can be compared
    tr._phase_d(vid, mel, spec, lens)
    tr._phase_g_pre(); tr._phase_g(); tr._phase_g2()
    torch.cuda.synchronize()
    names = [f"{mn}.{pn}" for mn in ("v_front", "gen", "post") for pn, _ in tr.mods[mn].named_parameters()]
    grads.append([p.grad.detach().clone() for p in tr.G.params])
    st = tr._st
    if run == 0:
        leaf0 = {k: st[k].grad.clone() for k in ("phon_leaf", "phon_g", "sent_g") if st.get(k) is not None and st[k].grad is not None}
    else:
        for k, v in leaf0.items():
            print("leaf grad", k, float((st[k].grad - v).norm() / (v.norm() + 1e-30)))
    del tr
    torch.cuda.empty_cache()
rows = []
for n, a, b in zip(names, grads[0], grads[1]):
    d = float((a - b).norm() / (a.norm() + 1e-30))
    rows.append((n, d, float(a.norm())))
bad = [r for r in rows if r[1] > 1e-5]
print(f"{len(bad)} of {len(rows)} parameters differ by more than 1e-5 between two identical runs")
for n, d, nm in rows:
    if (d > 1e-5 and "resnet" not in n) or n.endswith("decode.0.conv1.weight") or n.endswith("fc.weight") or n.endswith("layer4.1.conv2.weight"):
        print(f"{d:10.3e}  |g| {nm:10.3e}  {n}")
