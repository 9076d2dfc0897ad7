"""Timing of the reference algorithm as *stock PyTorch eager on the GPU* (cuDNN/cuBLAS, cudnn.benchmark=True as in
train.py:53-54) -- the denominator of the north star's ">= 10x reference-GPU-eager" target.  The reference sources
cannot travel to the GPU box, so this runs the oracle port (same torch ops, same schedule) with all tensors on cuda.
Not a test; run on the GPU box:   python tests/ref_gpu_eager_timing.py [B] [T] [fp32|bf16]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import GOLD  # noqa
from oracle import vca_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 75
mode = sys.argv[3] if len(sys.argv) > 3 else "fp32"
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda")
spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
sds = {}
for m in O.MODULES:
    sds[m] = {}
    for k, (shape, dt) in spec[m].items():
        t = O.det_tensor(m + "." + k, shape, getattr(torch, dt)).to(dev)
        if t.is_floating_point() and "running" not in k:
            t.requires_grad_(True)
        sds[m][k] = t
par = lambda ms: [{"params": [p for p in sds[m].values() if p.requires_grad]} for m in ms]  # noqa
g_opt = torch.optim.Adam(par(("v_front", "gen", "post")), lr=1e-4, weight_decay=1e-5, amsgrad=True)
d_opt = torch.optim.Adam(par(("dis1", "dis2", "dis3", "s_dis")), lr=1e-4, weight_decay=1e-5, amsgrad=True)
g = torch.Generator().manual_seed(1)
vid = torch.randn(B, 1, T, 112, 112, generator=g).to(dev)
mel = (torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1).to(dev)
sp = torch.rand(B, 1, 321, 4 * T, generator=g).to(dev)
lens = [T] * B
# av_attention builds its mask on the CPU: patch arange/as_tensor onto the device for this timing
_ar, _at = torch.arange, torch.as_tensor
torch.arange = lambda *a, **k: _ar(*a, **{**k, "device": dev})
torch.as_tensor = lambda *a, **k: _at(*a, **{**k, "device": dev})


def step():
    noise = torch.randn(B, 128, 20, T).to(dev)          # host RNG + H2D exactly as generator.py:248
    if mode == "bf16":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.train_step_with_adam(sds, dict(mel=mel, spec=sp, vid=vid, vid_len=lens), noise, g_opt, d_opt)
    return O.train_step_with_adam(sds, dict(mel=mel, spec=sp, vid=vid, vid_len=lens), noise, g_opt, d_opt)


for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
print(json.dumps({"what": "reference algorithm, PyTorch eager on GPU (oracle port on cuda)", "mode": mode, "B": B, "T": T,
                  "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "tf32": torch.backends.cudnn.allow_tf32}))
