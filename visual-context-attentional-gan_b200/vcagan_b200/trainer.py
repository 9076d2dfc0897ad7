"""The G+D training step of the reference (train.py:166-237; LRS variant train_LRS.py:179-243) on the B200-native
modules, with flat parameter/gradient buffers, a fused Adam(amsgrad) kernel and one-process-per-GPU data
parallelism (sum all-reduce of the flat gradient buffers over NCCL/NVLink; BatchNorm statistics stay per replica,
which is exactly what the reference's nn.DataParallel computes -- SURVEY.md 2.2 / 8e).

Schedule kept from the reference: D phase (real + R1 on the unconditional logits with create_graph, fake on
detached mels, sync loss with NON-detached phon so the visual front-end CNN receives its gradient), D optimizer
step, then the G phase against the *updated* discriminators; v_front gradients of both phases are summed before
the single G optimizer step.  Exact savings taken: D weight gradients are not
computed in the G phase (the reference computes and discards them, train.py:235-236).
"""
import math
from typing import Dict, Iterable, List, Optional

import os

import torch
import torch.nn as nn

from . import dp
from . import models as M
from . import ops
from ._lib import lib

DENORM_SCALE = -math.log(1e-5) / 2.0   # |d denormalize / d mel|, src/data/vid_aud_grid.py:238-240


class FlatGroup:
    """Re-homes the parameters of several modules into one flat fp32 buffer (and their .grad into another) so the
    optimizer is one kernel launch and the data-parallel exchange is one (bucketed) all-reduce."""

    def __init__(self, modules: Iterable[nn.Module]):
        self.params: List[nn.Parameter] = [p for m in modules for p in m.parameters()]
        self.epoch = [0]   # bumped by FusedAdam.step; part of the packed-weight cache tag (ops._packed)
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        # 16-byte align every segment
        self.offsets, off = [], 0
        for n in self.sizes:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o, n in zip(self.params, self.offsets, self.sizes):
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view(p.shape)
            p.grad = self.grad[o:o + n].view(p.shape)
            p._vca_epoch = self.epoch
            p._vca_flat = True   # ops._grad_sink: backward kernels may accumulate straight into p.grad
        # Tap-major gradient slabs for the conv filters (ops._conv_wgrad_raw): the tcgen05 wgrad kernels add [taps][Cout][Cin]
        # tiles with TMA reduce boxes instead of scattering 4-byte atomics over the [Cout][Cin][taps] parameter layout;
        # flush_slabs() folds them into .grad (one launch) and leaves them zeroed.
        self.slab = torch.zeros(off, dtype=torch.float32, device=dev)
        self._slab_jobs = []     # (param index, row of the job table without the CTA offset)
        for i, (p, o, n) in enumerate(zip(self.params, self.offsets, self.sizes)):
            if p.dim() == 4 and p.shape[2] * p.shape[3] > 1 and p.shape[1] % 4 == 0:
                taps = p.shape[2] * p.shape[3]
                p._vca_slab = self.slab[o:o + n].view(taps, p.shape[0], p.shape[1])
                self._slab_jobs.append((i, [self.grad.data_ptr() + 4 * o, self.slab.data_ptr() + 4 * o, 0, p.shape[0], p.shape[1], taps]))
        self._slab_tables = {}

    def _slab_table(self, lo: int, hi: int):
        key = (lo, hi)
        tab = self._slab_tables.get(key)
        if tab is None:
            rows, cta = [], 0
            for i, row in self._slab_jobs:
                if lo <= self.offsets[i] < hi:
                    rows.append(row + [cta, 0])
                    cta += lib().query("vca_unslab_job_ctas", row[3], row[4], row[5])
            tab = (torch.tensor(rows, dtype=torch.int64, device=self.grad.device) if rows else None, len(rows), cta,
                   max((r[5] for r in rows), default=1))
            self._slab_tables[key] = tab
        return tab

    def flush_slabs(self, lo: int = 0, hi: Optional[int] = None):
        """grad += slab (transposed into the parameter layout), slab = 0, for the parameters whose flat offset lies in
        [lo, hi) -- the same element ranges FusedAdam.step / the data-parallel all-reduce take.  Call it on a stream that
        is ordered after every backward kernel of those parameters (Trainer: right after _join_branches)."""
        tab = self._slab_table(lo, self.numel if hi is None else hi)
        if tab[0] is not None:
            lib().call("vca_grad_unslab_batched", tab[0], tab[1], tab[2], tab[3])

    def zero_grad(self):
        self.grad.zero_()
        for p, o, n in zip(self.params, self.offsets, self.sizes):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + n].view(p.shape)


class FusedAdam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay (coupled L2), amsgrad) as ONE CUDA kernel over
    the flat buffer (vca_adam_step).  train.py:82-83 uses amsgrad=True, train_LRS.py:97-98 amsgrad=False."""

    def __init__(self, group: FlatGroup, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, amsgrad=True):
        self.g, self.lr, self.betas, self.eps, self.wd = group, lr, betas, eps, weight_decay
        self.m = torch.zeros_like(group.flat)
        self.v = torch.zeros_like(group.flat)
        self.vmax = torch.zeros_like(group.flat) if amsgrad else None
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=group.flat.device)   # step count lives on the device
        # ... and so does the learning rate: the kernel reads it, so an lr decay (train.py:85-86, MultiStepLR([500, 800],
        # 0.1)) takes effect in captured CUDA graphs too -- set_lr() writes one float, no re-capture
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=group.flat.device)

    @property
    def t(self):
        return int(self.t_dev.item())

    def set_lr(self, lr: float):
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    @property
    def param_groups(self):
        """torch.optim-style view for lr schedulers / logging (train.py:165 prints param_groups[0]['lr'])."""
        return [{"lr": self.lr, "params": self.g.params}]

    def zero_grad(self):
        self.g.zero_grad()

    def step(self, grad_scale: float = 1.0, lo: int = 0, hi: Optional[int] = None, bump: bool = True, invalidate: bool = True):
        """One Adam step over flat[lo:hi] (default: the whole group).  A step may be taken in several slices -- the first
        with bump=True (it advances the step counter), the others with bump=False.  invalidate=False leaves the group's
        packed-weight cache tag alone: for an EARLY slice whose weights nobody reads before the last slice (which
        invalidates) -- otherwise the other slice's still-valid packs would be redone mid-backward."""
        hi = self.g.numel if hi is None else hi
        vmax = None if self.vmax is None else self.vmax[lo:hi]
        lib().call("vca_adam_step_dev", self.g.flat[lo:hi], self.g.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], vmax, hi - lo, self.lr_dev,
                   self.betas[0], self.betas[1], self.eps, self.wd, self.t_dev, grad_scale, 1 if bump else 0)
        ops._train_touch[0] += 1   # ... and the cache of BatchNorm-folded inference weights (ops.conv_folded)
        if invalidate:
            self.g.epoch[0] += 1   # the kernel wrote through raw pointers: invalidate the packed-weight cache of this group


def bilinear_down(mel: torch.Tensor, factor: int) -> torch.Tensor:
    """F.interpolate(mel, scale_factor=1/factor, mode='bilinear') for factor 2 or 4 on (B,1,H,W) (train.py:170-171):
    align_corners=False makes it the mean of the 2x2 pixels at rows/cols {2d,2d+1} (x0.5) or {4d+1,4d+2} (x0.25)."""
    x = mel.permute(0, 2, 3, 1)                       # channels-last view (C = 1)
    if factor == 2:
        y = ops.avg_pool2(x.contiguous())
    elif factor == 4:
        y = ops.avg_pool2(x[:, 1:, 1:, :].contiguous())[:, ::2, ::2, :]
    else:
        raise ValueError(factor)
    return y.permute(0, 3, 1, 2).contiguous()


class Trainer:
    def __init__(self, precision="bf16", lr=1e-4, weight_decay=1e-5, lrs=False, temp=1.0, device="cuda",
                 state: Optional[Dict[str, dict]] = None, dropout=True, process_group=None):
        ops.set_precision(precision)
        self.lrs = lrs
        self.merge_vfront_backward = True   # exact (up to fp re-association); False = the reference's two traversals
        self.parallel_branches = True       # the 3 discriminators + sync discriminator run as concurrent stream branches
        # capture(): the whole step as ONE CUDA graph.  Single-GPU only: with the NCCL all-reduces captured as graph nodes a
        # 2-GPU run hung on this pool (torch 2.11 / NCCL 2.28.9; killed by its timeout, not investigated further), so
        # data-parallel runs capture one graph per phase and issue the exchanges between them (set after self.world below).
        self.single_graph = True
        self.overlap_gru = True             # the sentence GRU runs on a side stream underneath the generator's first six blocks
        self._gru_stream = None
        self.batched_pack = True            # one weight re-pack launch per optimizer step (ops.PackPlan) instead of ~75
        self._pack_g = self._pack_d = None  # built after the first step (the pack cache then lists what the step uses)
        self._branch_streams = None
        self._br_used = []                  # branch streams with work since the last join
        self.device = torch.device(device)
        self.mods = dict(v_front=M.Visual_front(1), gen=M.Decoder(), post=M.Postnet(), dis1=M.Discriminator(phase='1'),
                         dis2=M.Discriminator(phase='2'), dis3=M.Discriminator(phase='3'), s_dis=M.sync_Discriminator(temp))
        if state is not None:
            for k, m in self.mods.items():
                m.load_state_dict(state[k])
        for m in self.mods.values():
            m.to(self.device).train()
        if not dropout:
            self.mods["v_front"].dropout.p = 0.0
            self.mods["v_front"].sentence_encoder.dropout = 0.0
        self.G = FlatGroup([self.mods[k] for k in ("v_front", "gen", "post")])
        self.D = FlatGroup([self.mods[k] for k in ("dis1", "dis2", "dis3", "s_dis")])
        self.g_opt = FusedAdam(self.G, lr, weight_decay=weight_decay, amsgrad=not lrs)
        self.d_opt = FusedAdam(self.D, lr, weight_decay=weight_decay, amsgrad=not lrs)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self.single_graph = self.world == 1
        # VCA_NATIVE_COMM=1: the gradient exchange goes through the library's own NCCL communicator (vca_comm_init /
        # vca_allreduce_bucket) instead of torch.distributed's -- same NCCL, same sums (tests/test_gpu_dp2.py)
        self.native_comm = dp.NativeComm(process_group, self.device) if (self.world > 1 and os.environ.get("VCA_NATIVE_COMM") == "1") else None
        if self.world > 1:
            self._sync_replicas()
        # Data-parallel runs split the G backward where the generator's last gradient is written: the all-reduce of the
        # gen + post gradients (62 % of the G buffer) then runs on the comm stream underneath the visual front-end's
        # backward.  Single-GPU runs keep the backward in one piece (no all-reduce to hide, one graph less).
        self.split_g_backward = self.world > 1
        # Single GPU: the optimizer passes (Adam over 36 B / parameter, weight re-pack -- pure HBM streaming) run on a side
        # stream underneath compute-bound work that does not touch what they write: the D optimizer + re-pack under the
        # Postnet forward / reconstruction losses, Adam on the gen + post slice under the visual front-end's backward (which
        # needs the G backward split where data-parallel runs split it).  Set before the first step / capture().
        # Measured (1 x B200, same box, B = 32): 47.50 ms with, 47.48 ms without -- the HBM-streaming optimizer kernels take from
        # the concurrent compute kernels what they gain.  Negative result: off by default (VCA_OVERLAP_OPT=1 turns it on).
        self.overlap_opt = self.world == 1 and os.environ.get("VCA_OVERLAP_OPT", "0") == "1"
        if os.environ.get("VCA_SPLIT_G") == "1":
            self.split_g_backward = True
        if self.overlap_opt:
            self.split_g_backward = True
        self._opt_stream = None
        self._d_done = self._ga_done = False
        n_vf = sum(1 for _ in self.mods["v_front"].parameters())
        self._vf_params = self.G.params[:n_vf]
        # job tables of every slab flush the schedules use, built now (a host-to-device copy is illegal inside a graph capture)
        vf_numel = self.G.offsets[n_vf]
        for grp, rng in ((self.D, (0, self.D.numel)), (self.G, (0, self.G.numel)), (self.G, (vf_numel, self.G.numel)), (self.G, (0, vf_numel))):
            grp._slab_table(*rng)
        self._genpost_params = self.G.params[n_vf:]
        self._vf_numel = self.G.offsets[n_vf]             # v_front gradients are G.grad[:_vf_numel]
        # weight / bias gradients that accumulate straight into the flat .grad buffer overlap the dgrad chain
        ops.cfg.param_grad_streams = tuple(torch.cuda.Stream(device=self.device) for _ in range(2)) if self.parallel_branches else ()
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # the branch streams are intentional

    # -- data-parallel exchange ---------------------------------------------------------------------------------
    def _sync_replicas(self):
        """Data-parallel start-up: every replica takes rank 0's parameters and BatchNorm buffers -- only
        gradients are exchanged afterwards, so replicas that start different stay different -- and every rank moves to
        its own Philox sub-stream, so dropout masks and generator noise differ across the shards of a global batch."""
        import torch.distributed as dist
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None and self.pg is not dist.group.WORLD else 0
        dp.broadcast_flat(self.G.flat, src, self.pg)
        dp.broadcast_flat(self.D.flat, src, self.pg)
        bufs = [b for m in self.mods.values() for b in m.buffers()]
        fl = [b for b in bufs if b.is_floating_point()]
        if fl:
            flat = torch.cat([b.reshape(-1).float() for b in fl])
            dp.broadcast_flat(flat, src, self.pg)
            o = 0
            for b in fl:
                b.copy_(flat[o:o + b.numel()].view(b.shape)); o += b.numel()
        for b in bufs:
            if not b.is_floating_point():
                dist.broadcast(b, src, group=self.pg)
        ops.set_rng_rank(dist.get_rank(self.pg))

    def replicas_in_sync(self) -> bool:
        """True when the parameter checksums of all ranks are bit-identical (they must be after any number of steps)."""
        if self.world == 1:
            return True
        import torch.distributed as dist
        cs = torch.stack([self.G.flat.double().sum(), self.D.flat.double().sum()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.pg)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.pg)
        return bool(torch.equal(lo, hi))

    def _allreduce(self, group: FlatGroup, bucket_elems: int = 8 << 20, lo: int = 0, hi: Optional[int] = None,
                   wait: bool = True):
        """Sum-all-reduce group.grad[lo:hi] on the comm stream, ordered after the current stream.  wait=False leaves it
        running underneath whatever the current stream does next; the next waiting call (or an explicit
        `cur.wait_stream(comm_stream)`) joins it."""
        if self.world == 1:
            return
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            if self.native_comm is not None:       # the C ABI's own communicator (vca_allreduce_bucket)
                self.native_comm.allreduce_flat(group.grad[lo:group.numel if hi is None else hi], bucket_elems)
            else:
                dp.allreduce_flat(group.grad[lo:group.numel if hi is None else hi], self.pg, bucket_elems)
            done = torch.cuda.Event()
            done.record()
        if wait:
            cur.wait_stream(self.comm_stream)
        return done

    def _branches(self, fns):
        """Run independent sub-graphs concurrently: fns[0] on the current stream, the others on side streams that fork
        from / join back into it (in a CUDA-graph capture these become parallel branches).  The discriminators share no
        parameters and only read common inputs, and autograd replays each branch's backward on the stream its forward
        ran on, so the small-grid kernels of the low-resolution discriminators fill the SMs that the tails of the
        full-resolution one leave idle -- forward and backward."""
        if not self.parallel_branches or len(fns) < 2:
            return [f() for f in fns]
        assert len(fns) == 4
        self._ensure_branch_streams()
        cur = torch.cuda.current_stream()
        outs, used = [None] * len(fns), []
        for i in range(1, len(fns)):
            st = self._branch_streams[(i - 1) % len(self._branch_streams)]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs[i] = fns[i]()
            used.append(st)
        self._br_used.extend(used)          # their backward kernels will run there too: joined by _join_branches
        outs[0] = fns[0]()
        for st in used:
            cur.wait_stream(st)
        return outs

    def _ensure_branch_streams(self):
        if self._branch_streams is None:
            self._branch_streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(3)]   # every fork uses all three

    def _fork(self, fns, stream_idx):
        """Start fns[k] on branch stream stream_idx[k] (ordered after the current stream) WITHOUT joining: the current
        stream carries on with independent work; `_join_branches` (or the next `_branches` fork) orders it back."""
        if not self.parallel_branches:
            return [f() for f in fns]
        self._ensure_branch_streams()
        cur = torch.cuda.current_stream()
        outs = []
        for f, k in zip(fns, stream_idx):
            st = self._branch_streams[k]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(f())
            self._br_used.append(st)
        return outs

    def _join_branches(self):
        """After a backward pass: the branch backward kernels (incl. the wgrad kernels that accumulate straight into
        the flat .grad buffer, which autograd's own leaf-stream bookkeeping does not see) ran on the side streams."""
        cur = torch.cuda.current_stream()
        # only streams that took part since the last join: inside a graph capture, waiting on a stream that is not part of
        # the capture would tie the graph to work outside it and invalidate the capture
        for st in dict.fromkeys(self._br_used):
            cur.wait_stream(st)
        self._br_used.clear()
        for st in ops.cfg._pg_used:         # parameter-gradient side streams with work since the last join
            cur.wait_stream(st)
        ops.cfg._pg_used.clear()

    # -- one step -------------------------------------------------------------------------------------------------
    def step(self, vid, mel, spec, vid_len, noise=None):
        """vid (B,1,T,112,112), mel (B,1,80,4T), spec (B,1,321,4T) device fp32; vid_len int32 device tensor or list."""
        phases = [lambda: self._phase_d(vid, mel, spec, vid_len, noise), self._phase_g_pre, self._phase_g, self._phase_g2,
                  self._phase_end_a, self._phase_end_b]
        out = self._run_schedule(lambda i: phases[i]())
        self._build_pack_plans()
        return out

    def _run_schedule(self, run, replaying=False):
        """The step as six phases with the data-parallel exchanges between them; `run(i)` executes phase i (eagerly, or by
        replaying the CUDA graph it was captured into).  Every all-reduce is issued on the comm stream and only waited for
        where its result is needed, so each one runs underneath independent work:
          0 D phase | all-reduce D grads ~ 1 Postnet forward + reconstruction losses (no discriminator involved)
          2 D optimizer, G phase down to the generator's leaves | all-reduce gen + post grads ~ 3 visual front-end backward
          | all-reduce v_front grads ~ 4 Adam on gen + post | 5 Adam on v_front, weight re-pack.
        Single-GPU runs have no exchanges; phases 1+2 and 4+5 then share a graph (3 graphs per step)."""
        cur = torch.cuda.current_stream()
        multi = self.world > 1
        ovl = self.overlap_opt and not multi and not replaying and self.split_g_backward
        if ovl and self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=self.device)
        side = self._opt_stream
        run(0)
        ev = self._allreduce(self.D, wait=False) if multi else None
        if ovl:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._d_update()
            self._d_done = True
        run(1)
        if multi:
            cur.wait_event(ev)
        if ovl:
            cur.wait_stream(side)
        run(2)
        if self.split_g_backward:
            e1 = self._allreduce(self.G, lo=self._vf_numel, wait=False) if multi else None      # gen + post
            if ovl:
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    # (no cache invalidation yet: the visual front-end's backward is about to use ITS packed weights, which
                    #  share the group's tag; _phase_end_b's slice invalidates before the batched re-pack)
                    self.g_opt.step(1.0 / self.world, lo=self._vf_numel, bump=True, invalidate=False)
                self._ga_done = True
            run(3)
            e2 = self._allreduce(self.G, hi=self._vf_numel, wait=False) if multi else None      # v_front
            if multi:
                cur.wait_event(e1)
            if ovl:
                cur.wait_stream(side)
            run(4)
            if multi:
                cur.wait_event(e2)
        else:
            if multi:
                self._allreduce(self.G)
            run(4)
        return run(5)

    def _phase_groups(self):
        """phases captured together into one CUDA graph"""
        if self.world > 1:
            return [[0], [1], [2], [3], [4], [5]] if self.split_g_backward else [[0], [1], [2], [4, 5]]
        return [[0], [1, 2, 3], [4, 5]] if self.split_g_backward else [[0], [1, 2], [4, 5]]

    def _phase_d(self, vid, mel, spec, vid_len, noise=None):
        """forward of v_front + generator, the whole D phase and its backward (train.py:168-210)."""
        m = self.mods
        v_front, gen = m["v_front"], m["gen"]
        dis = (m["dis1"], m["dis2"], m["dis3"])
        s_dis = m["s_dis"]
        gen.fixed_noise = noise
        ops.cfg.deferred_counters = []      # BatchNorm.num_batches_tracked of the whole step: one foreach add in _phase_g
        self.G.zero_grad()                                                # train.py:168
        mel1, mel2 = bilinear_down(mel, 4), bilinear_down(mel, 2)          # train.py:170-171
        T = vid.size(2)
        reals = [t.detach().requires_grad_(True) for t in (mel1, mel2, mel)]

        def real_early(i):
            """trunk + unconditional head + R1 penalty of the real pass (train.py:176-194): needs no generator output and
            no sentence embedding, so it runs on a branch stream underneath the visual front-end / generator forward"""
            def run():
                h = dis[i].features(reals[i], T)
                u = dis[i].uncond_head(h)
                gr = torch.autograd.grad(u.sum(), reals[i], create_graph=True)[0]     # R1 (train.py:188-194)
                return h, u, ops.sum_sq(gr, 1.0 / gr.size(0))
            return run
        early = self._fork([real_early(2), real_early(1), real_early(0)], [2, 0, 1])
        early = {2: early[0], 1: early[1], 0: early[2]}
        phon = v_front.features(vid)
        # The 2-layer bi-GRU is 4 x T strictly sequential steps on a few clusters (1.25 ms with most SMs idle); the
        # generator's stem (decode x3 + g1 x3, the heaviest convolutions) needs only `phon`.  Run the GRU on its own
        # stream underneath the stem; autograd replays its backward there too, underneath the stem's backward.
        cur = torch.cuda.current_stream()
        side = self.overlap_gru and self.parallel_branches
        if side:
            if self._gru_stream is None:
                self._gru_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._gru_stream.wait_stream(cur)
            with torch.cuda.stream(self._gru_stream):
                sent = v_front.sentence(phon)
            self._br_used.append(self._gru_stream)
        else:
            sent = v_front.sentence(phon)
        if self.split_g_backward:
            # the generator sees detached leaves: its backward stops there (_phase_g) and the visual front-end's
            # backward is a second autograd call fed with their gradients (_phase_g2)
            assert self.merge_vfront_backward
            phon_g = phon.detach().requires_grad_(True)
            h = gen.stem(phon_g)
            if side:
                cur.wait_stream(self._gru_stream)
            sent_g = sent.detach().requires_grad_(True)
            g = gen.tail(sent_g, h, vid_len)                               # g1, g2, g3
        else:
            phon_g = sent_g = None
            h = gen.stem(phon)
            if side:
                cur.wait_stream(self._gru_stream)
            g = gen.tail(sent, h, vid_len)
        gen.fixed_noise = None
        assert phon.size(1) == T
        sdet = sent.detach()
        g_det = [x.detach() for x in g]
        self._join_branches()

        def d_branch(i):
            """conditional head of the real pass, then discriminator i on its fake mel (train.py:176-200)"""
            def run():
                h, u, pen = early[i]
                c = dis[i].cond_head(h, sdet, T)
                uf_, cf_ = dis[i](g_det[i], sdet, T)
                return u, c, pen, uf_, cf_
            return run
        # phon is NOT detached in the reference (train.py:186): the sync loss sends a gradient into the visual
        # front-end CNN during the D backward, and the G backward traverses that CNN a second time.  The CNN weights
        # do not change in between, so we take d(dis_loss)/d(phon) here on a detached leaf and inject it into the
        # single G-phase traversal below -- the same sum of the two gradients, one CNN backward instead of two.
        phon_leaf = phon.detach().requires_grad_(True) if self.merge_vfront_backward else phon
        res = self._branches([d_branch(2), d_branch(1), d_branch(0), lambda: s_dis(phon_leaf, reals[2]).mean()])
        sync_loss = res[3]
        by_d = [res[2], res[1], res[0]]                                    # back to dis1, dis2, dis3 order
        ur, cr, gp = [r[0] for r in by_d], [r[1] for r in by_d], [r[2] for r in by_d]
        uf, cf = [r[3] for r in by_d], [r[4] for r in by_d]
        real_loss = sum(M.gan_loss(x, True) for x in ur + cr) / 3 + sum(gp) / 3
        fake_loss = sum(M.gan_loss(x, False) for x in uf + cf) / 3
        dis_loss = real_loss + fake_loss + (0.5 if self.lrs else 1.0) * sync_loss
        self.d_opt.zero_grad()
        if self.merge_vfront_backward:
            dis_loss.backward(inputs=self.D.params + [phon_leaf])
        else:
            dis_loss.backward(retain_graph=True, inputs=self.D.params + self._vf_cnn_params())
        self._join_branches()
        self.D.flush_slabs()         # tap-major wgrad slabs -> .grad (the all-reduce / Adam that follow read .grad)
        self._st = dict(mel=mel, mel1=mel1, mel2=mel2, spec=spec, phon=phon, phon_leaf=phon_leaf, sdet=sdet, g=g, T=T,
                        sent=sent, phon_g=phon_g, sent_g=sent_g,
                        out=dict(dis_loss=dis_loss.detach(), sync_loss=sync_loss.detach(), real_loss=real_loss.detach(),
                                 fake_loss=fake_loss.detach(), grad_pen=torch.stack([t.detach() for t in gp])))

    def _phase_g_pre(self):
        """The part of the G phase that involves no discriminator (train.py:215, 226-229): Postnet forward and the four L1
        reconstruction terms.  In data-parallel runs it executes underneath the all-reduce of the D gradients."""
        st, m = self._st, self.mods
        g = st["g"]
        gs = m["post"](g[2])
        k = 1.0 if self.lrs else DENORM_SCALE                              # GRID: L1 on de-normalised mels
        recon = (ops.l1_mean(g[0], st["mel1"], k) + ops.l1_mean(g[1], st["mel2"], k) + ops.l1_mean(g[2], st["mel"], k)) / 3 \
            + ops.l1_mean(gs, st["spec"])
        st["gs"], st["recon"] = gs, recon

    def _d_update(self):
        """D optimizer step on the (exchanged) gradients + one batched re-pack of the weights it changed"""
        self.d_opt.step(1.0 / self.world)
        if self._pack_d is not None:
            self._pack_d.run()

    def _phase_g(self):
        """D optimizer step, then the G phase against the updated discriminators and its backward (train.py:211-236)."""
        st, m = self._st, self.mods
        dis = (m["dis1"], m["dis2"], m["dis3"])
        g, sdet, T, phon = st["g"], st["sdet"], st["T"], st["phon"]
        if self._d_done:
            self._d_done = False             # already stepped on the optimizer side stream (underneath _phase_g_pre)
        else:
            self._d_update()
        gs, recon = st["gs"], st["recon"]
        res = self._branches([lambda: dis[2](g[2], sdet, T), lambda: dis[1](g[1], sdet, T), lambda: dis[0](g[0], sdet, T),
                              lambda: m["s_dis"](phon.detach(), g[2], True).mean()])
        ug, cg = [res[2][0], res[1][0], res[0][0]], [res[2][1], res[1][1], res[0][1]]
        g_sync = res[3]
        g_adv = sum(M.gan_loss(x, True) for x in ug + cg) / 3
        gen_loss = g_adv + g_sync + 50.0 * recon
        if self._gru_stream is not None and not self.split_g_backward:
            self._br_used.append(self._gru_stream)       # the GRU's backward kernels will run there: join it afterwards
        # D weight grads are skipped (the reference computes and discards them, train.py:235-236)
        if self.split_g_backward:
            torch.autograd.backward([gen_loss], inputs=self._genpost_params + [st["phon_g"], st["sent_g"]])
        elif self.merge_vfront_backward:
            torch.autograd.backward([gen_loss, phon], [None, st["phon_leaf"].grad], inputs=self.G.params)
        else:
            gen_loss.backward(inputs=self.G.params)
        self._join_branches()
        if self.split_g_backward:
            self.G.flush_slabs(lo=self._vf_numel)       # gen + post: reduced / stepped first
        else:
            self.G.flush_slabs()
        ops.flush_deferred_counters()
        st["out"].update(gen_loss=gen_loss.detach(), g_sync=g_sync.detach(), recon=recon.detach(), g1=g[0].detach(),
                         g2=g[1].detach(), g3=g[2].detach(), gs=gs.detach())

    def _phase_g2(self):
        """Second half of a split G backward: the visual front-end, fed with d(gen_loss)/d(phon, sent) from the
        generator's leaves and d(dis_loss)/d(phon) from the D phase (autograd sums the two roots on phon)."""
        st = self._st
        if not self.split_g_backward:
            return
        if self._gru_stream is not None:
            self._br_used.append(self._gru_stream)
        torch.autograd.backward([st["phon"], st["phon"], st["sent"]],
                                [st["phon_g"].grad, st["phon_leaf"].grad, st["sent_g"].grad], inputs=self._vf_params)
        self._join_branches()
        self.G.flush_slabs(hi=self._vf_numel)

    def _phase_end_a(self):
        """G optimizer: with a split backward the gen + post slice first (its gradients were reduced long ago), so that the
        v_front slice's all-reduce finishes underneath it; otherwise the whole group."""
        if self._ga_done:
            self._ga_done = False            # gen + post slice already stepped on the side stream (underneath _phase_g2)
        elif self.split_g_backward:
            self.g_opt.step(1.0 / self.world, lo=self._vf_numel, bump=True)
        else:
            self.g_opt.step(1.0 / self.world)

    def _phase_end_b(self):
        if self.split_g_backward:
            self.g_opt.step(1.0 / self.world, lo=0, hi=self._vf_numel, bump=False)
        if self._pack_g is not None:
            self._pack_g.run()
        out, self._st = self._st["out"], None
        return out

    def _build_pack_plans(self):
        """After a full step the pack cache holds every (view of a) parameter the step packs: from now on they are
        refreshed by ONE launch per optimizer step into the same persistent slabs."""
        if not self.batched_pack or self._pack_g is not None or ops.cfg.dtype != torch.bfloat16:
            return
        self._pack_g, self._pack_d = ops.PackPlan(self.G.params), ops.PackPlan(self.D.params)
        self._pack_g.run(); self._pack_d.run()

    # -- CUDA-graph replay of the step ----------------------------------------------------------------------------
    def capture(self, vid, mel, spec, vid_len, warmup=3, noise=None):
        """Capture the step into CUDA graphs sharing one memory pool (3 on one GPU: D phase | G phase | G optimizer; 6 in
        data-parallel runs, see _run_schedule); the gradient all-reduces run between them.  Everything that changes from step to step lives in device memory
        (Adam step counters, Philox stream positions, BN buffers), so `replay` needs no host-side state.
        `vid_len` must be an int32 device tensor.  Inputs are copied into static buffers on every replay."""
        assert torch.is_tensor(vid_len) and vid_len.is_cuda, "vid_len must be a device tensor for graph capture"
        self._sin = [t.clone() for t in (vid, mel, spec)] + [vid_len.to(torch.int32).clone()]
        self._snoise = None if noise is None else noise.to(self.device).clone()   # parity tests only
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(*self._sin, noise=self._snoise)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.clear_pack_cache(keep=(self._pack_g, self._pack_d))   # every other weight (re)pack must be recorded inside the graphs
        pool = torch.cuda.graph_pool_handle()
        groups = [[0, 1, 2, 3, 4, 5]] if self.single_graph else self._phase_groups()
        self._graphs = [torch.cuda.CUDAGraph() for _ in groups]
        self._graph_of = {grp[0]: i for i, grp in enumerate(groups)}      # first phase of a group -> its graph
        n0 = lib().launches
        # The critical path (main chain + discriminator branches) is captured on high-priority streams, the
        # parameter-gradient side streams keep the default (lowest) priority: when both have CTAs pending, the SMs
        # go to the chain the step is waiting for.
        cap = torch.cuda.Stream(device=self.device, priority=-1)
        phases = [lambda: self._phase_d(*self._sin, noise=self._snoise), self._phase_g_pre, self._phase_g, self._phase_g2,
                  self._phase_end_a, self._phase_end_b]
        if self.single_graph:
            # One graph for the whole step: no drain between phases (the tail of one phase overlaps the head of the next).
            with torch.cuda.graph(self._graphs[0], pool=pool, stream=cap):
                self._sout = self._run_schedule(lambda i: phases[i]())
        else:
            for gi, grp in enumerate(groups):
                with torch.cuda.graph(self._graphs[gi], pool=pool, stream=cap):
                    for i in grp:
                        r = phases[i]()
                        if i == 5:
                            self._sout = r
        self.launches_per_step = lib().launches - n0
        return self

    def replay(self, vid=None, mel=None, spec=None, vid_len=None):
        """One training step from the captured graphs; new inputs (host-pinned or device) are copied into the static buffers."""
        for dst, src in zip(self._sin, (vid, mel, spec, vid_len)):
            if src is not None:
                dst.copy_(src, non_blocking=True)
        if self.single_graph:
            self._graphs[0].replay()
            return self._sout

        def run(i):
            gi = self._graph_of.get(i)
            if gi is not None:
                self._graphs[gi].replay()
            return self._sout if i == 5 else None
        return self._run_schedule(run, replaying=True)

    # -- pipelined input feed: the host->device copy of step i+1 runs on a copy stream underneath step i ------------
    def stage_inputs(self, vid, mel, spec):
        """Start the asynchronous copy of (pinned host or device) inputs into the staging buffers."""
        if getattr(self, "_stage", None) is None:
            self._stage = [torch.empty_like(t) for t in self._sin[:3]]
            self._copy_stream = torch.cuda.Stream(device=self.device)
        self._copy_stream.wait_stream(torch.cuda.current_stream())   # the previous step's D2D reads of the staging buffers
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage, (vid, mel, spec)):
                dst.copy_(src, non_blocking=True)

    def replay_prefetched(self, nxt=None):
        """One captured step on the inputs staged by `stage_inputs` / the previous call; `nxt` = (vid, mel, spec) of the
        NEXT step, whose host->device copy overlaps this step's compute.  Every step's inputs still cross PCIe exactly
        once; only the wait for them moves off the critical path."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self._copy_stream)
        for dst, src in zip(self._sin[:3], self._stage):
            dst.copy_(src, non_blocking=True)                         # device-to-device, 136 MB at HBM speed
        if nxt is not None:
            self.stage_inputs(*nxt)
        return self.replay()

    def _vf_cnn_params(self):
        vf = self.mods["v_front"]
        return list(vf.frontend.parameters()) + list(vf.resnet.parameters())

    def state_dicts(self):
        """checkpoint layout of train.py:303-309"""
        return {f"{k}_state_dict": m.state_dict() for k, m in self.mods.items()}
