#!/usr/bin/env python
"""bench.py -- G+D train samples/sec of the VCA-GAN hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--frames T]

One "step" = one full G+D training step (train.py:166-237: visual front-end, generator, three multi-scale
discriminators with R1, sync discriminator, Postnet, both fused Adam(amsgrad) updates) on one batch of synthetic
GRID-shape clips (config[1] of BASELINE.json: batch 32 per GPU, 75 frames of 112x112 lips -> 80x300 mel, bf16).
N > 1 is launched by torchrun: one process per GPU, batch 32 per GPU (weak scaling; N = 8 is the global-batch-256
config[2]), sum all-reduce of the flat gradient buffers over NCCL.

Prints ONE JSON line on rank 0 (see the keys below).  `value` is timed with inputs resident in HBM; `e2e` re-times
the same step through the public API with pinned HOST inputs (H2D inside the timed region) and a D2H read of the
losses.  `--impl reference` times the reference algorithm's CPU implementation (the oracle port -- the reference's
own sources cannot travel to the GPU box) on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "visual-context-attentional-gan_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

GFLOP_PER_SAMPLE = {40: 455.0, 50: 567.8, 75: 852.0, 250: 2850.7}   # BASELINE.md section 3 (reference step as written)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append([c.strip() for c in o])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v == "Active":
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth(B, T, seed):
    """Synthetic GRID-shape batch: unit-scale frames, mel in [-1,1], spec >= 0, full lengths."""
    g = torch.Generator().manual_seed(seed)
    vid = torch.randn(B, 1, T, 112, 112, generator=g)
    mel = torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1
    spec = torch.rand(B, 1, 321, 4 * T, generator=g)
    return vid, mel, spec


# -----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference step on the host cores
# -----------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(T, steps, warmup, B=2):
    from oracle import vca_oracle as O
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_spec.json")))
    torch.set_num_threads(os.cpu_count() or 1)
    sds = {}
    for m in O.MODULES:
        sds[m] = {}
        for k, (shape, dt) in spec[m].items():
            t = O.det_tensor(m + "." + k, shape, getattr(torch, dt))
            if t.is_floating_point() and "running" not in k:
                t.requires_grad_(True)
            sds[m][k] = t
    par = lambda ms: [{"params": [p for p in sds[m].values() if p.requires_grad]} for m in ms]  # noqa: E731
    g_opt = torch.optim.Adam(par(("v_front", "gen", "post")), lr=1e-4, weight_decay=1e-5, amsgrad=True)
    d_opt = torch.optim.Adam(par(("dis1", "dis2", "dis3", "s_dis")), lr=1e-4, weight_decay=1e-5, amsgrad=True)
    vid, mel, sp = synth(B, T, 1)
    noise = torch.randn(B, 128, 20, T)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step_with_adam(sds, dict(mel=mel, spec=sp, vid=vid, vid_len=[T] * B), noise, g_opt, d_opt)
        times.append(time.perf_counter() - t0)
    tt = times[warmup:]
    return B * len(tt) / sum(tt), sum(tt) / len(tt), B


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = args.frames
    # bounded sample: B=2 clips per step keeps a (warmup+steps) run within minutes on the box's host cores
    sps, sec, B = cpu_reference_steps(T, args.steps, args.warmup, B=2)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "G+D train samples/sec (GRID 3s clips)", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GRID G+D train step, T={T} frames, 112x112 lips -> 80x{4 * T} mel (BASELINE config[1])",
                   "per_step_sample": f"B={B} clips per CPU step (bounded sample of the B=32 workload)"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} G+D steps of B={B}, T={T}, fp32, oracle port of train.py:166-237, torch CPU {cores} threads"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# -----------------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import vcagan_b200 as V
    from vcagan_b200.trainer import Trainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    assert V.lib().query("vca_device_ok") == 1, "bench needs an sm_100 GPU (no fallback path exists)"
    B, T = args.batch, args.frames
    torch.manual_seed(1)
    V.manual_seed(1 + rank)
    tr = Trainer(precision=args.precision, dropout=True, device=dev, process_group=pg)
    if world > 1:   # identical replicas: broadcast rank 0's weights once
        import torch.distributed as dist
        dist.broadcast(tr.G.flat, 0); dist.broadcast(tr.D.flat, 0)
    vid_h, mel_h, spec_h = [t.pin_memory() for t in synth(B, T, 100 + rank)]
    vid, mel, spec = vid_h.to(dev), mel_h.to(dev), spec_h.to(dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, host_inputs):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n):
            l2_flush.zero_()
            ev[i][0].record()
            if host_inputs:
                out = run_host(vid_h, mel_h, spec_h)                        # H2D of this step's inputs from pinned memory
                _ = torch.stack([out["gen_loss"], out["dis_loss"]]).cpu()   # D2H read of the step's result
            else:
                out = run_resident()
            ev[i][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / n, out

    if args.no_graph:
        for _ in range(max(args.warmup, 3)):
            tr.step(vid, mel, spec, lens)
        run_resident = lambda: tr.step(vid, mel, spec, lens)                       # noqa: E731
        run_host = lambda v, m_, s_: tr.step(v.to(dev, non_blocking=True), m_.to(dev, non_blocking=True),  # noqa: E731
                                             s_.to(dev, non_blocking=True), lens)
    else:   # whole step replayed from CUDA graphs (no host launch overhead; optimizer/RNG state is device resident)
        tr.capture(vid, mel, spec, lens, warmup=max(args.warmup, 3))
        run_resident = lambda: tr.replay()                                         # noqa: E731
        run_host = lambda v, m_, s_: tr.replay(v, m_, s_)                          # noqa: E731
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = V.lib().launches
    ms, out = timed(args.steps, False)
    launches = (V.lib().launches - n0) // args.steps if args.no_graph else tr.launches_per_step
    barrier()
    clocks = sampler.summary()
    n_e2e = max(2, min(args.steps, 5))
    if args.no_graph:
        ms_e2e, out2 = timed(n_e2e, True)
    else:
        # End to end through the pipelined input feed (Trainer.stage_inputs / replay_prefetched): ONE timed region over
        # n_e2e steps that contains the pinned-host -> device copy of every one of those steps' inputs (step 1's up
        # front, step i+1's on a copy stream underneath step i) and a device -> host read of the losses every step.
        # No L2 flush here: each step's fresh 136 MB of inputs alone exceed the 126 MB L2.
        tr.stage_inputs(vid_h, mel_h, spec_h)            # untimed warm-up of the feed path (allocates the staging buffers)
        for i in range(2):
            out2 = tr.replay_prefetched((vid_h, mel_h, spec_h) if i == 0 else None)
            _ = torch.stack([out2["gen_loss"], out2["dis_loss"]]).cpu()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.stage_inputs(vid_h, mel_h, spec_h)
        for i in range(n_e2e):
            out2 = tr.replay_prefetched((vid_h, mel_h, spec_h) if i + 1 < n_e2e else None)
            _ = torch.stack([out2["gen_loss"], out2["dis_loss"]]).cpu()
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / n_e2e
    barrier()
    assert torch.isfinite(out["gen_loss"]).item() and torch.isfinite(out["dis_loss"]).item(), "non-finite loss"
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    # data-parallel invariant: every rank applied the same averaged gradients to the same weights, so the replicas must
    # still be bit-identical after all those steps (checksum of both flat parameter buffers, min == max over ranks)
    in_sync = None
    if world > 1:
        cs = torch.stack([tr.G.flat.double().sum(), tr.D.flat.double().sum()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
        assert in_sync, f"replicas diverged: parameter checksums {lo.tolist()} .. {hi.tolist()}"

    # dominant kernel family: the tcgen05 implicit-GEMM conv (fwd + dgrad share one kernel); timed live per launch
    roof = None
    # every rank runs the instrumented step (it contains the gradient all-reduces); only rank 0 reports it
    # (serialised: the concurrent stream branches are switched off so that each launch is timed alone on its stream)
    tr.parallel_branches = False
    V.ops.cfg.param_grad_streams = ()
    prof = V.lib().profile_step(lambda: tr.step(vid, mel, spec, lens))
    barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank == 0:
        pk = peaks()
        fam ={k: v for k, v in prof.items() if k.startswith("vca_conv_") and (k.endswith("_tc") or k.endswith("_tc_ws"))}
        flops = sum(v["flops"] for v in fam.values()); tms = sum(v["ms"] for v in fam.values())
        n_l = sum(v["n"] for v in fam.values())
        total_ms = sum(v["ms"] for v in prof.values())
        if tms > 0:
            ach = flops / (tms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "conv_tc_fwd_kernel+conv_tc_wgrad_kernel (tcgen05 implicit GEMM)",
                    "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"], "traffic": None,
                    "launches_per_step": n_l, "share_of_kernel_time": tms / total_ms, "peak_source": pk["src"] + " sustained"}
        top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({k: v for k, v in top}, open(os.path.join(ROOT, "gpurun_out", "bench_kernel_breakdown.json"), "w"), indent=1)
    if rank != 0:
        return
    sps = world * B / (ms * 1e-3)
    gf = GFLOP_PER_SAMPLE.get(T)
    pk = peaks()
    cpu_sps, cpu_sec, cpu_b = cpu_reference_steps(T, 1, 0, B=2) if not args.no_cpu_baseline else (None, None, 2)
    line = {
        "metric": "G+D train samples/sec (GRID 3s clips)", "value": sps, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"GRID G+D train step (BASELINE config[1]), batch {B}/GPU, T={T}, 112x112 lips -> 80x{4 * T} mel",
                   "global_batch": world * B, "parallelism": f"dp{world}", "replicas_in_sync": in_sync, "l2": "256 MiB flush buffer written between timed steps",
                   "launch": "eager" if args.no_graph else (
                       "3 CUDA graphs per step (D phase | G phase | G optimizer)" if world == 1 else
                       "4 CUDA graphs per step (D phase | G phase to the generator's leaves | visual front-end backward | G "
                       "optimizer); NCCL all-reduce of D grads between 1-2, of gen+post grads underneath graph 3, of v_front "
                       "grads before graph 4"),
                   "e2e_feed": "per-step blocking copy" if args.no_graph else
                               "one timed region over all e2e steps; step i+1's H2D copy overlaps step i on a copy stream",
                   "step_tensor_roofline_frac": (sps / world * gf * 1e9 / (pk["tf_sust"] * 1e12)) if gf else None,
                   "algorithmic_gflop_per_sample": gf},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": (vid_h.numel() + mel_h.numel() + spec_h.numel()) * 4, "d2h_bytes_per_step": 8},
        "roofline": roof,
        "cpu_baseline": None if cpu_sps is None else {
            "value": cpu_sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"1 G+D step of B={cpu_b}, T={T}, fp32, oracle port of train.py:166-237 on torch CPU ({cpu_sec:.1f} s)"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=75)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
