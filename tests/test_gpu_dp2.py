"""Multi-GPU correctness on real GPUs: the gradients a 2-rank data-parallel step ends up with
(NCCL sum all-reduce of the flat buffers, 1/world folded into Adam) must equal the mean of the gradients of two
SINGLE-GPU steps, one per shard -- BatchNorm statistics are per replica in the reference's nn.DataParallel and here, so
per-shard execution is the exact oracle.  Also: a replica constructed with different weights is overwritten by rank 0's
at Trainer start-up, and the replicas are bit-identical after the step.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import make_state, rel_l2, GOLD, ROOT
from oracle import vca_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_gradients_equal_per_shard_mean(tmp_path, precision):
    import vcagan_b200 as V
    from vcagan_b200.trainer import Trainer
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dp2_worker as W
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dp2_worker.py"), str(tmp_path), precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ranks = [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(2)]
    assert ranks[0]["in_sync"] and ranks[1]["in_sync"]
    assert ranks[0]["Gw"] == ranks[1]["Gw"] and ranks[0]["Dw"] == ranks[1]["Dw"]          # bit-identical replicas after the step
    assert torch.equal(ranks[0]["G"], ranks[1]["G"]) and torch.equal(ranks[0]["D"], ranks[1]["D"])
    spec = json.load(open(os.path.join(GOLD, "state_spec.json")))
    try:
        # The per-shard oracle: two single-GPU trainers, one per shard.  The D phase (forward, D losses, D backward) of a
        # shard depends on nothing outside the shard.  The G phase runs against the discriminators UPDATED with the
        # rank-averaged D gradient (train.py:211 precedes :217), so the oracle hands both trainers the mean of their D
        # gradients before their G phase -- which is all the all-reduce does.
        trs, outs = [], []
        for shard in range(2):
            state = {m: make_state(spec, m) for m in O.MODULES}
            tr = Trainer(precision=precision, state=state, dropout=False)
            vid, mel, sp, noise, lens = W.shard_inputs(shard)
            tr._phase_d(vid.cuda(), mel.cuda(), sp.cuda(), lens, noise)
            trs.append(tr)
        torch.cuda.synchronize()
        d_mean = (trs[0].D.grad + trs[1].D.grad) / 2.0
        e = rel_l2(W.sample(d_mean), ranks[0]["D"] / 2.0)
        print(f"{precision}: 2-rank all-reduced D gradients vs mean of per-shard single-GPU gradients: rel L2 {e:.3e}")
        # fp32: measured 2e-8 (D) / 1e-7 (G).  bf16: the forward of the tcgen05 path is bit-reproducible since round 2 (split-K
        # into per-split slabs added in order, BatchNorm-statistic partials reduced in a fixed order inside the CTA), and what
        # is left in the gradients is the fp32 summation order of the wgrad reduce-adds: measured 2e-8 (D) / 1e-7 (G) between
        # two identical runs (tools/g_noise_probe.py; it was 8e-2 / 3.6e-1 while shared-slab atomics flipped bf16 roundings in
        # front of the train-mode BatchNorms), so the bf16 exchange is held to 1e-4 too
        tol = 2e-5 if precision == "fp32" else 1e-4
        assert e < tol, e
        # The G phase runs against discriminators stepped with the EXCHANGED gradient.  Adam's first step moves every weight
        # by +-lr whatever the gradient's size, so the bf16 noise in d_mean (e above: ~1e-1) would flip thousands of update
        # signs and hand the oracle a different discriminator than the ranks trained against (measured: G mismatch 0.4).
        # The oracle therefore applies the ranks' own reduced D gradient, which the D check above has just validated.
        d_used = torch.load(os.path.join(tmp_path, "d_grad_sum.pt")).cuda() / 2.0
        assert rel_l2(d_used, d_mean) < tol
        # bf16 yard-stick, measured live: a third single-GPU trainer repeats shard 0 with the same D gradient (run-to-run
        # figure of the G gradient; ~1e-7 now that the forward is bit-reproducible -- printed for the record, the bound is tol)
        g_noise = None
        if precision != "fp32":
            tr3 = Trainer(precision=precision, state={m: make_state(spec, m) for m in O.MODULES}, dropout=False)
            vid, mel, sp, noise, lens = W.shard_inputs(0)
            tr3._phase_d(vid.cuda(), mel.cuda(), sp.cuda(), lens, noise)
            trs.append(tr3)
        for tr in trs:
            tr.D.grad.copy_(d_used)
            tr._phase_g_pre(); tr._phase_g(); tr._phase_g2(); tr._phase_end_a()
            outs.append(tr._phase_end_b())
        torch.cuda.synchronize()
        if precision != "fp32":
            g_noise = rel_l2(W.sample(trs[2].G.grad), W.sample(trs[0].G.grad))
            print(f"{precision}: run-to-run noise of the single-GPU G gradient (same shard, same D gradient): {g_noise:.3e}")
        for i in range(2):      # each rank's losses are its own shard's losses
            for k, v in outs[i].items():
                if torch.is_tensor(v) and v.numel() == 1:
                    assert abs(ranks[i]["losses"][k] - float(v)) <= (2e-5 if precision == "fp32" else 3e-2) * max(1.0, abs(float(v))), (i, k)
        g_mean = (trs[0].G.grad + trs[1].G.grad) / 2.0
        e = rel_l2(W.sample(g_mean), ranks[0]["G"] / 2.0)
        print(f"{precision}: 2-rank all-reduced G gradients vs mean of per-shard single-GPU gradients: rel L2 {e:.3e}")
        # by slice: the visual front-end (train-mode BatchNorm at B = 2: ill-conditioned) and generator + Postnet
        n_s, cut = ranks[0]["G"].numel(), int(round(trs[0]._vf_numel / trs[0].G.numel * ranks[0]["G"].numel()))
        sm, sr = W.sample(g_mean), ranks[0]["G"] / 2.0
        e_vf, e_gp = rel_l2(sm[:cut], sr[:cut]), rel_l2(sm[cut:], sr[cut:])
        print(f"{precision}:   v_front slice {e_vf:.3e}  (|g| {float(sr[:cut].norm()):.3e}), gen+post slice {e_gp:.3e} (|g| {float(sr[cut:].norm()):.3e})")
        assert e < (tol if g_noise is None else max(tol, 2.0 * g_noise)), (e, g_noise)
        # ... and therefore the same weights after the step (both single-GPU oracles applied g_mean? no: each applied its own
        # shard's G gradient -- only the D weights are comparable)
        if precision == "fp32":
            assert abs(float(trs[0].D.flat.double().sum()) - ranks[0]["Dw"]) <= 1e-6 * abs(ranks[0]["Dw"]) + 1e-3
    finally:
        V.set_precision("fp32")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_native_comm_matches_torch_distributed(tmp_path):
    """include/vcagan.h's NCCL helpers (vca_comm_unique_id / vca_comm_init / vca_allreduce_bucket): bucketed in-place
    sum all-reduce of fp32 / bf16 buffers is bit-identical to torch.distributed.all_reduce on the same data (two ranks: the
    sum of two addends is order-independent), and a Trainer step whose gradient exchange runs through it (VCA_NATIVE_COMM=1)
    ends with the same gradients as the torch.distributed run of test_two_rank_gradients_equal_per_shard_mean."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "comm_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ranks = [torch.load(os.path.join(tmp_path, f"comm_rank{i}.pt")) for i in range(2)]
    for rk in ranks:
        assert rk["world"] == 2 and rk["f32"] and rk["bf16"] and rk["tiny"] and rk["in_sync"]
    assert torch.equal(ranks[0]["G"], ranks[1]["G"]) and torch.equal(ranks[0]["D"], ranks[1]["D"])
    # against the torch.distributed exchange of the same two shards
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29544", os.path.join(ROOT, "tests", "dp2_worker.py"), str(tmp_path), "fp32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(os.path.join(tmp_path, "rank0.pt"))
    eg, ed = rel_l2(ranks[0]["G"], ref["G"]), rel_l2(ranks[0]["D"], ref["D"])
    print("native-comm step vs torch.distributed step: G", eg, "D", ed)
    assert eg < 2e-5 and ed < 2e-5
