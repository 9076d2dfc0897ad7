#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of the training step (B = 32, T = 75, bf16), through torch.profiler (CUPTI):
for every kernel its name, stream, start and duration AS IT RAN INSIDE THE GRAPH, i.e. with the concurrent stream
branches, not serialised as under ncu.  Writes gpurun_out/step_trace.json (a compact list) and prints a summary:
busy time per kernel family, SM-idle gaps on the union of all streams, and the longest kernels.
    python tools/step_trace.py [B] [T] [tag]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-context-attentional-gan_b200")); sys.path.insert(0, ROOT)
import torch
import vcagan_b200 as V
from vcagan_b200.trainer import Trainer
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 75
tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
dev = torch.device("cuda")
torch.manual_seed(1); V.manual_seed(1)
tr = Trainer(precision="bf16", dropout=True, device=dev)
g = torch.Generator().manual_seed(3)
vid = torch.randn(B, 1, T, 112, 112, generator=g).to(dev)
mel = (torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1).to(dev)
spec = torch.rand(B, 1, 321, 4 * T, generator=g).to(dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
tr.capture(vid, mel, spec, lens, warmup=3)
for _ in range(3):
    tr.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.replay()
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        ev.append((e.name, int(getattr(e, "device_index", 0)), e.time_range.start, e.time_range.end))
# stream ids are only in the chrome trace: export and re-read
path = os.path.join(ROOT, "gpurun_out", f"step_trace_{tag}_chrome.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
tr_js = json.load(open(path))
ks = [(x["name"], x["args"].get("stream", -1), x["ts"], x["dur"]) for x in tr_js["traceEvents"]
      if x.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in x]
ks.sort(key=lambda k: k[2])
t0 = ks[0][2]
import re


def short(n):   # torch's templated kernels: keep the functor that says what they do
    if "at::" in n:
        m = re.findall(r"(\w+Functor\w*|\w*[Cc]opy\w*|CatArray\w*|\w+_kernel_cuda\w*|reduce_kernel|multi_tensor_apply\w*)", n)
        return ("torch:" + "/".join(dict.fromkeys(m[:3])))[:60] if m else n[:60]
    return n[:60]


out = [(short(n), s, round(ts - t0, 3), round(d, 3)) for n, s, ts, d in ks]
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"step_trace_{tag}.json"), "w"))
os.remove(path)
end = max(ts + d for _, _, ts, d in out)
print(f"{len(out)} kernels, span {end / 1e3:.2f} ms")
