#!/usr/bin/env python
"""Where the (eager, concurrent-branch) training step spends its wall time on the main stream: CUDA events at the
phase boundaries of Trainer.step, incl. gradient-arrival hooks inside the two backward passes.
   python tools/phase_times.py [B] [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200.trainer import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 75
dev = torch.device("cuda")
torch.manual_seed(0); V.manual_seed(0)
tr = Trainer(precision="bf16", dropout=True, device=dev)
g = torch.Generator().manual_seed(3)
vid = torch.randn(B, 1, T, 112, 112, generator=g).to(dev)
mel = torch.randn(B, 1, 80, 4 * T, generator=g).to(dev)
spec = torch.rand(B, 1, 321, 4 * T, generator=g).to(dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    tr.step(vid, mel, spec, lens)
torch.cuda.synchronize()

marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def hooked(mod, name_after):
    orig = mod.forward

    def fwd(*a, **k):
        out = orig(*a, **k)
        mark(name_after)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        for i, o in enumerate(outs):
            if torch.is_tensor(o) and o.requires_grad:
                o.register_hook(lambda gr, n=f"grad reached {name_after.split()[0]} out[{i}]": (mark(n), None)[1])
        return out
    mod.forward = fwd
    return orig


o1 = hooked(tr.mods["v_front"], "v_front fwd done")
o2 = hooked(tr.mods["gen"], "gen fwd done")
o3 = hooked(tr.mods["post"], "post fwd done")
pd, pg, pe = tr._phase_d, tr._phase_g, tr._phase_end
tr._phase_d = lambda *a, **k: (mark("step start"), pd(*a, **k), mark("D phase done (fwd + D backward)"))[1]
tr._phase_g = lambda *a, **k: (pg(*a, **k), mark("G phase done (D opt, D fwd on fakes, G backward)"))[0]
tr._phase_end = lambda *a, **k: (pe(*a, **k), mark("G optimizer done"))[0]
for rep in range(2):
    marks.clear()
    tr.step(vid, mel, spec, lens)
    torch.cuda.synchronize()
t0 = marks[0][1]
prev = 0.0
print(f"B={B} T={T}: eager step, concurrent branches on; cumulative ms on the stream each event was recorded on")
for name, e in marks:
    t = t0.elapsed_time(e)
    print(f"  {t:8.2f} ms  (+{t - prev:6.2f})  {name}")
    prev = t

# ---- the same step from the three captured graphs (no host launch overhead): device time per graph
tr._phase_d, tr._phase_g, tr._phase_end = pd, pg, pe
tr.mods["v_front"].forward, tr.mods["gen"].forward, tr.mods["post"].forward = o1, o2, o3
tr.capture(vid, mel, spec, lens)
for _ in range(3):
    tr.replay()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = [0.0, 0.0, 0.0]
for _ in range(5):
    ev[0].record(); tr._graphs[0].replay(); ev[1].record(); tr._graphs[1].replay(); ev[2].record(); tr._graphs[2].replay(); ev[3].record()
    torch.cuda.synchronize()
    for i in range(3):
        acc[i] += ev[i].elapsed_time(ev[i + 1]) / 5
print(f"CUDA-graph replay: D-phase graph {acc[0]:.2f} ms | G-phase graph {acc[1]:.2f} ms | G optimizer graph {acc[2]:.2f} ms | total {sum(acc):.2f} ms")
