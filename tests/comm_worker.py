"""Worker of tests/test_gpu_dp2.py::test_native_comm_*: the library's own NCCL communicator (vca_comm_init / vca_allreduce_bucket)
against torch.distributed's all-reduce on the same buffers, and one Trainer step whose gradient exchange runs through it."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "visual-context-attentional-gan_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    out_dir = sys.argv[1]
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vcagan_b200 import dp
    comm = dp.NativeComm(dist.group.WORLD, dev)
    res = {"world": comm.world}
    g = torch.Generator().manual_seed(100 + rank)
    for name, dt, n in (("f32", torch.float32, (20 << 20) + 12345), ("bf16", torch.bfloat16, 1 << 20), ("tiny", torch.float32, 3)):
        x = torch.randn(n, generator=g).to(dev).to(dt)
        a, b = x.clone(), x.clone()
        dist.all_reduce(a)
        comm.allreduce_flat(b, bucket_elems=8 << 20)
        torch.cuda.synchronize()
        res[name] = bool(torch.equal(a, b))
        res[name + "_sum"] = float(b.double().sum())
    # a full data-parallel step with the exchange on the native communicator
    os.environ["VCA_NATIVE_COMM"] = "1"
    from conftest import make_state, GOLD
    from oracle import vca_oracle as O
    from vcagan_b200.trainer import Trainer
    import dp2_worker as W
    spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
    tr = Trainer(precision="fp32", state={m: make_state(spec, m) for m in O.MODULES}, dropout=False, device=dev, process_group=dist.group.WORLD)
    assert tr.native_comm is not None
    vid, mel, sp, noise, lens = W.shard_inputs(rank)
    tr.step(vid.to(dev), mel.to(dev), sp.to(dev), lens, noise=noise)
    torch.cuda.synchronize()
    res.update(G=W.sample(tr.G.grad), D=W.sample(tr.D.grad), in_sync=tr.replicas_in_sync())
    torch.save(res, os.path.join(out_dir, f"comm_rank{rank}.pt"))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
