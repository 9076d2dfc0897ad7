"""Host-side data-parallel logic on CPU with world_size-2 gloo: batch sharding, bucketed sum
all-reduce of a flat gradient buffer (ragged tail bucket included), weight broadcast, and the semantic identity the
design relies on -- averaging per-rank gradients of per-replica-BatchNorm shards equals what the reference's
nn.DataParallel computes (checked with the oracle's Postnet, which contains a train-mode BatchNorm)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG, make_state

WORLD = 2


def _worker(rank, port, tmp):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from vcagan_b200 import dp
        import json
        from oracle import vca_oracle as O
        # ---- sharding
        lo, hi = dp.shard_bounds(256, WORLD, rank)
        assert (lo, hi) == (rank * 128, (rank + 1) * 128)
        with pytest.raises(ValueError):
            dp.shard_bounds(255, WORLD, rank)
        assert dp.bucket_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
        # ---- bucketed all-reduce with a ragged tail, sync and async
        n = 1000 + 37
        flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
        dp.allreduce_flat(flat, None, bucket_elems=256)
        assert torch.equal(flat, torch.arange(n, dtype=torch.float32) * 3)
        flat = torch.full((n,), float(rank + 1))
        for w in dp.allreduce_flat(flat, None, bucket_elems=300, async_op=True):
            w.wait()
        assert torch.equal(flat, torch.full((n,), 3.0))
        # ---- the split G all-reduce of the trainer (gen + post slice first, v_front slice later) == one all-reduce of the
        #      whole buffer: in-place reduction of two disjoint views of the flat gradient buffer
        whole = torch.arange(n, dtype=torch.float32) * (rank + 1)
        parts = whole.clone()
        cut = 413                                   # not a bucket multiple, like Trainer._vf_numel
        dp.allreduce_flat(parts[cut:], None, bucket_elems=256)
        assert torch.equal(parts[:cut], whole[:cut])          # the other slice is untouched in between
        dp.allreduce_flat(parts[:cut], None, bucket_elems=256)
        dp.allreduce_flat(whole, None, bucket_elems=256)
        assert torch.equal(parts, whole)
        # ---- broadcast makes replicas identical
        w0 = torch.randn(50) if rank == 0 else torch.zeros(50)
        dp.broadcast_flat(w0, 0)
        ref = [torch.zeros(50) for _ in range(WORLD)]
        dist.all_gather(ref, w0)
        assert torch.equal(ref[0], ref[1])
        # ---- DP semantics: mean over ranks of shard gradients (per-replica BN) == single process doing each shard
        spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_spec.json")))
        g = torch.Generator().manual_seed(3)
        mel = torch.rand(4, 1, 80, 24, generator=g) * 2 - 1
        tgt = torch.rand(4, 1, 321, 24, generator=g)

        def shard_grads(lo, hi):
            sd = make_state(spec, "post", requires_grad=True)
            loss = torch.nn.functional.l1_loss(O.postnet(sd, mel[lo:hi], True), tgt[lo:hi])
            loss.backward()
            return torch.cat([v.grad.reshape(-1) for v in sd.values() if v.requires_grad])
        lo, hi = dp.shard_bounds(4, WORLD, rank)
        mine = shard_grads(lo, hi)
        dp.allreduce_flat(mine, None, bucket_elems=1 << 16)
        mine /= WORLD
        expect = (shard_grads(0, 2) + shard_grads(2, 4)) / 2
        assert torch.allclose(mine, expect, rtol=1e-5, atol=1e-7)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_dp_host_logic_gloo_world2(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(WORLD))
