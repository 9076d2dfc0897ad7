"""Worker of tests/test_gpu_dp2.py: one rank of a 2-GPU data-parallel Trainer step (launched by torchrun).
Writes the rank's all-reduced gradient buffers (sampled), its losses and post-step parameter checksums to <out>/rank<r>.pt."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "visual-context-attentional-gan_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import make_state, GOLD  # noqa: E402
from oracle import vca_oracle as O  # noqa: E402


def shard_inputs(rank, B=2, T=20):
    g = torch.Generator().manual_seed(4321 + rank)
    vid = torch.randn(B, 1, T, 112, 112, generator=g)
    mel = torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1
    spec = torch.rand(B, 1, 321, 4 * T, generator=g)
    noise = torch.randn(B, 128, 20, T, generator=g)
    lens = [T, T - 3 - rank]
    return vid, mel, spec, noise, lens


def sample(flat, n=200000):
    idx = torch.linspace(0, flat.numel() - 1, n, dtype=torch.float64).round().long().to(flat.device)
    return flat[idx].double().cpu()


def main():
    out_dir, precision = sys.argv[1], sys.argv[2]
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vcagan_b200.trainer import Trainer
    spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
    state = {m: make_state(spec, m) for m in O.MODULES}
    if rank == 1:      # a replica that starts DIFFERENT must be overwritten by rank 0's weights at Trainer start-up
        for sd in state.values():
            for k, v in sd.items():
                if v.is_floating_point():
                    sd[k] = v + 0.01
    tr = Trainer(precision=precision, state=state, dropout=False, device=dev, process_group=dist.group.WORLD)
    vid, mel, sp, noise, lens = shard_inputs(rank)
    out = tr.step(vid.to(dev), mel.to(dev), sp.to(dev), lens, noise=noise)
    torch.cuda.synchronize()
    res = dict(G=sample(tr.G.grad), D=sample(tr.D.grad), Gw=float(tr.G.flat.double().sum()), Dw=float(tr.D.flat.double().sum()),
               losses={k: float(v) for k, v in out.items() if torch.is_tensor(v) and v.numel() == 1}, in_sync=tr.replicas_in_sync())
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    if rank == 0:      # the exchanged D gradient in full: the per-shard oracle applies exactly this one (see the test)
        torch.save(tr.D.grad.cpu(), os.path.join(out_dir, "d_grad_sum.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
