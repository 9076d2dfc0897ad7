#!/usr/bin/env python
"""Kernel launch census of one eager training step (torch.profiler, CUDA activities): how many launches are ours and
how many are torch data-movement / autograd glue, with and without fused gradient accumulation.
   python tools/count_kernels.py [B] [T]"""
import collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200"))
import torch
import vcagan_b200 as V
from vcagan_b200.trainer import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 25
dev = torch.device("cuda")
torch.manual_seed(0)
V.manual_seed(0)
tr = Trainer(precision="bf16", dropout=True, device=dev)
g = torch.Generator().manual_seed(3)
vid = torch.randn(B, 1, T, 112, 112, generator=g).to(dev)
mel = torch.randn(B, 1, 80, 4 * T, generator=g).to(dev)
spec = torch.rand(B, 1, 321, 4 * T, generator=g).to(dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(2):
    tr.step(vid, mel, spec, lens)
torch.cuda.synchronize()


def census(tag):
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        tr.step(vid, mel, spec, lens)
        torch.cuda.synchronize()
    cnt, tim = collections.Counter(), collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            n = e.name
            if "at::" in n:
                m = re.search(r"(FillFunctor|CUDAFunctor_add|CUDAFunctorOnSelf_add|direct_copy|CatArray|MulFunctor|reduce_kernel)", n)
                n = "torch:" + (m.group(1) if m else n[:40])
            else:
                n = re.sub(r"<.*", "", n).replace("void ", "").split("::")[-1][:40]
            cnt[n] += 1; tim[n] += e.device_time
    tot = sum(cnt.values())
    print(f"== {tag}: {tot} kernel launches, {sum(tim.values()) / 1e3:.2f} ms of kernel time")
    for n, c in cnt.most_common(14):
        print(f"   {c:5d} x {n:42s} {tim[n] / 1e3:8.3f} ms")
    return cnt


a = census("fused gradient accumulation ON")
V.ops.cfg.fuse_grad_accum = False
tr.step(vid, mel, spec, lens)
b = census("fused gradient accumulation OFF")
