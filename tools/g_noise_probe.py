#!/usr/bin/env python
"""How reproducible is the bf16 G-phase gradient at the B = 2, T = 20 test case?  Three single-GPU trainers on the SAME
shard and the same D gradient: merged backward twice (run-to-run noise) and the split backward of the data-parallel
schedule (tests/test_gpu_dp2.py compares a 2-rank split run with merged single-GPU runs)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "visual-context-attentional-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import make_state, rel_l2, GOLD
from oracle import vca_oracle as O
import dp2_worker as W
from vcagan_b200.trainer import Trainer

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
serial = len(sys.argv) > 2 and sys.argv[2] == "serial"
import vcagan_b200 as V
spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
trs = []
for split in (False, False, True):
    tr = Trainer(precision=prec, state={m: make_state(spec, m) for m in O.MODULES}, dropout=False)
    tr.split_g_backward = split
    if serial:
        tr.parallel_branches = False; tr.overlap_gru = False
        V.ops.cfg.param_grad_streams = ()
    vid, mel, sp, noise, lens = W.shard_inputs(0)
    tr._phase_d(vid.cuda(), mel.cuda(), sp.cuda(), lens, noise)
    trs.append(tr)
torch.cuda.synchronize()
print("D grads: run-to-run", rel_l2(trs[1].D.grad, trs[0].D.grad), " split-vs-merged", rel_l2(trs[2].D.grad, trs[0].D.grad))
d = trs[0].D.grad.clone()
for tr in trs:
    tr.D.grad.copy_(d)
    tr._phase_g_pre(); tr._phase_g(); tr._phase_g2()
torch.cuda.synchronize()
cut = trs[0]._vf_numel
for name, a in (("run-to-run", trs[1]), ("split-vs-merged", trs[2])):
    print(f"G grads {name}: all {rel_l2(a.G.grad, trs[0].G.grad):.3e}  v_front {rel_l2(a.G.grad[:cut], trs[0].G.grad[:cut]):.3e}  "
          f"gen+post {rel_l2(a.G.grad[cut:], trs[0].G.grad[cut:]):.3e}")
for k in ("phon", "sent"):
    print(k, "forward 1-vs-0", rel_l2(trs[1]._st[k].float(), trs[0]._st[k].float()), "2-vs-1", rel_l2(trs[2]._st[k].float(), trs[1]._st[k].float()))
# the visual front-end alone, twice on the same module and input (eval of determinism without any trainer state)
vf = trs[0].mods["v_front"]
vid = W.shard_inputs(0)[0].cuda()
with torch.no_grad():
    a = vf(vid) if not isinstance(vf(vid), tuple) else vf(vid)[0]
    b = vf(vid) if not isinstance(vf(vid), tuple) else vf(vid)[0]
    c = trs[1].mods["v_front"](vid); c = c[0] if isinstance(c, tuple) else c
print("v_front twice on trainer 0:", rel_l2(b.float(), a.float()), " trainer 1 vs trainer 0 module:", rel_l2(c.float(), a.float()))
w0 = dict(trs[0].mods["v_front"].named_parameters()); w1 = dict(trs[1].mods["v_front"].named_parameters())
print("max param diff 0 vs 1 AFTER the G phases (no optimizer step was run):", max(float((w0[k] - w1[k]).abs().max()) for k in w0))
b0 = dict(trs[0].mods["v_front"].named_buffers()); b1 = dict(trs[1].mods["v_front"].named_buffers())
print("max buffer diff:", max(float((b0[k].float() - b1[k].float()).abs().max()) for k in b0))
for k in ("g1", "g3", "gs"):
    print(k, "forward run-to-run", rel_l2(trs[1]._st["out"][k].float(), trs[0]._st["out"][k].float()), "split", rel_l2(trs[2]._st["out"][k].float(), trs[0]._st["out"][k].float()))
