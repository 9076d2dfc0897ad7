#!/usr/bin/env python
"""bench.py -- G+D train samples/sec of the VCA-GAN hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]
                    [--workload train|lrs|inference] [--scaling weak|strong] [--batch B] [--frames T]

workload train (default) = BASELINE config[1]/[2]: one "step" is one full G+D training step (train.py:166-237: visual
front-end, generator, three multi-scale discriminators with R1, sync discriminator, Postnet, both fused Adam(amsgrad)
updates) on one batch of synthetic GRID-shape clips (batch 32 per GPU, 75 frames of 112x112 lips -> 80x300 mel, bf16).
N > 1 is launched by torchrun: one process per GPU, sum all-reduce of the flat gradient buffers over NCCL.
`--scaling weak` keeps 32 clips per GPU (N = 8 is the global-batch-256 config[2]); `--scaling strong` fixes the global
batch (default 256) and gives each GPU 256/N clips.
workload lrs = BASELINE config[3]: train_LRS.py:179-243, clips padded to 250 frames, batch 16 per GPU, ragged vid_len.
workload inference = BASELINE config[4]: test.py:126-143, generator + flip TTA + Postnet + 60 Griffin-Lim iterations,
64 clips per GPU (replicas only: inference has no exchange step).

Prints ONE JSON line on rank 0.  `value` is timed with inputs resident in HBM; `e2e` re-times the same step through the
public API with pinned HOST inputs (H2D inside the timed region) and a D2H read of the result.
`--impl reference` times the UNMODIFIED reference modules (baseline/_ref, installed by baseline/install_reference.py)
through the stock step body (baseline/stock_step.py) on the box's host cores; `--impl reference-gpu` runs the same
modules as PyTorch eager on the B200 (cudnn.benchmark=True, train.py:53-54) -- the denominator of the north star's
">= 10x reference-GPU-eager" target, which the default line also carries as `gpu_eager_baseline`.
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
GFLOP_PER_CLIP_INFER = {75: 356.0}                                   # 2 x (v_front + gen) + post, T = 75
METRIC = {"train": "G+D train samples/sec (GRID 3s clips)", "lrs": "G+D train samples/sec (LRS 250-frame clips)",
          "inference": "test-time inference clips/sec (generator + flip TTA + Postnet + Griffin-Lim)"}


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


def synth(B, T, seed, lrs=False):
    """Synthetic batch: unit-scale frames, mel in [-1,1]; GRID: spec >= 0 (raw magnitudes,
    vid_aud_grid.py:146), full lengths; LRS: spec in [-1,1] (vid_aud_lrs2.py:176-178), ragged lengths in [T/2, T] with
    the frames behind a clip's end zeroed as collate_fn pads them (vid_aud_lrs2.py:203-233)."""
    g = torch.Generator().manual_seed(seed)
    vid = torch.randn(B, 1, T, 112, 112, generator=g)
    mel = torch.rand(B, 1, 80, 4 * T, generator=g) * 2 - 1
    spec = torch.rand(B, 1, 321, 4 * T, generator=g)
    if not lrs:
        return vid, mel, spec, torch.full((B,), T, dtype=torch.int64)
    spec = spec * 2 - 1
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T                                      # the batch is padded to its longest clip
    for i, n in enumerate(lens.tolist()):
        vid[i, :, n:] = 0.0
    return vid, mel, spec, lens


def workload_name(args, B, T):
    if args.workload == "inference":
        return f"GRID test-time inference (BASELINE config[4]), batch {B}/GPU, T={T}, flip TTA + Postnet + {args.gl_iters} Griffin-Lim iterations"
    if args.workload == "lrs":
        return f"LRS G+D train step (BASELINE config[3]), batch {B}/GPU, clips padded to T={T}, ragged vid_len, 112x112 lips -> 80x{4 * T} mel"
    return f"GRID G+D train step (BASELINE config[1]), batch {B}/GPU, T={T}, 112x112 lips -> 80x{4 * T} mel"


def shape_of(args, world):
    """(per-GPU batch, frames) of the run."""
    T = args.frames if args.frames else (250 if args.workload == "lrs" else 75)
    if args.batch:
        B = args.batch
    elif args.workload == "inference":
        B = 64
    elif args.workload == "lrs":
        B = 16
    elif args.scaling == "strong":
        assert args.global_batch % world == 0, "global batch must divide by the number of GPUs"
        B = args.global_batch // world
    else:
        B = 32
    return B, T


# -----------------------------------------------------------------------------------------------------------------
# reference arms: the unmodified reference modules (baseline/_ref) through the stock step body
# -----------------------------------------------------------------------------------------------------------------
def _librosa_stub():
    """src/data/stft.py:32 imports librosa (not installed, no network): the three functions it uses, per SURVEY 8(c)."""
    import types
    import numpy as np
    lib = types.ModuleType("librosa"); util = types.ModuleType("librosa.util"); filt = types.ModuleType("librosa.filters")
    util.pad_center = lambda data, size, **k: data
    util.tiny = lambda x: np.finfo(np.float32).tiny
    util.normalize = lambda x, norm=None, **k: x
    lib.util, lib.filters = util, filt
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt})


def reference_steps(workload, B, T, steps, warmup, device, autocast=False, gl_iters=60):
    """Time the unmodified reference on `device` ('cpu' or 'cuda').  -> (units/s, seconds/step, kind)."""
    from baseline import stock_step as S
    if not S.reference_available():
        return None
    cpu = device == "cpu"
    ns = S.import_reference(cpu_shim=cpu)
    torch.manual_seed(1)
    mods = S.build_modules(ns)
    lrs = workload == "lrs"
    if not cpu:
        torch.backends.cudnn.deterministic = False        # train.py:53-54
        torch.backends.cudnn.benchmark = True
        for m in mods.values():
            m.cuda()
    vid, mel, spec, lens = synth(B, T, 1, lrs)
    sync = (lambda: None) if cpu else torch.cuda.synchronize
    times = []
    if workload == "inference":
        _librosa_stub()
        saved = list(sys.path)
        sys.path[:] = [S.REF_DIR] + [q for q in saved if not os.path.isfile(os.path.join(q or ".", "src", "__init__.py"))]
        try:
            from src.data.stft import STFT
            from src.data.audio_processing import griffin_lim
        finally:
            sys.path[:] = saved
            for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
                sys.modules["_vca_ref2_" + k] = sys.modules.pop(k)
        stft = STFT(640, 160, 640)
        if not cpu:
            stft = stft.cuda()
        for m in mods.values():
            m.eval()
        v_front, gen, post = mods["v_front"], mods["gen"], mods["post"]

        def one():     # test.py:130-143
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and not cpu):
                v = vid.cuda()
                phon, sent = v_front(v)
                g3 = gen(sent, phon, lens)[2]
                phon, sent = v_front(v.flip(4))
                g3 = (g3 + gen(sent, phon, lens)[2]) / 2.
                gs = post(g3)
                wav = griffin_lim(gs.squeeze(1).float(), stft, gl_iters)
                return wav.cpu()
    else:
        for m in mods.values():
            m.train()
        g_opt, d_opt = S.build_optimizers(mods, lrs=lrs)

        def one():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and not cpu):
                return S.stock_train_step(mods, g_opt, d_opt, (mel.clone(), spec, vid, lens), ns.gan_loss, lrs=lrs)
    for i in range(warmup + steps):
        sync()
        t0 = time.perf_counter()
        one()
        sync()
        times.append(time.perf_counter() - t0)
    tt = times[warmup:]
    return B * len(tt) / sum(tt), sum(tt) / len(tt), "reference"


def oracle_port_steps(T, steps, warmup, B=2):
    """Fallback when baseline/_ref is absent: the oracle port of the reference step on the host cores."""
    from oracle import vca_oracle as O
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_spec.json")))
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
    vid, mel, sp, _ = synth(B, T, 1)
    noise = torch.randn(B, 128, 20, T)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step_with_adam(sds, dict(mel=mel, spec=sp, vid=vid, vid_len=[T] * B), noise, g_opt, d_opt)
        times.append(time.perf_counter() - t0)
    tt = times[warmup:]
    return B * len(tt) / sum(tt), sum(tt) / len(tt), "port"


def cpu_arm(workload, T, steps, warmup, gl_iters=60):
    """The reference's CPU implementation of the path on all host cores, on a bounded sample (B = 2 clips per step)."""
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    r = reference_steps(workload, B, T, steps, warmup, "cpu", gl_iters=gl_iters)
    if r is None:
        if workload != "train":
            raise RuntimeError("baseline/_ref is missing and the oracle port only covers the GRID train step")
        r = oracle_port_steps(T, steps, warmup, B)
    return r + (B,)


def base_config(args, world, B, T):
    return {"workload": workload_name(args, B, T), "global_batch": world * B,
            "parallelism": f"dp{world}" if args.workload != "inference" else f"replicas x{world}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B_ours, T = shape_of(args, world)
    sps, sec, kind, B = cpu_arm(args.workload, T, args.steps, args.warmup, args.gl_iters)
    cores = os.cpu_count() or 1
    what = "unmodified reference modules (baseline/_ref) through the stock step body" if kind == "reference" else \
        "oracle port of train.py:166-237 (baseline/_ref absent)"
    cfg = base_config(args, world, B_ours, T)
    line = {
        "impl": "reference", "metric": METRIC[args.workload], "value": sps, "unit": "clips/s" if args.workload == "inference" else "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps of B={B} clips (bounded sample of the B={B_ours} workload), T={T}, fp32, {what}, "
                                   f"torch CPU {cores} threads"},
        "e2e": {"value": sps, "unit": "clips/s" if args.workload == "inference" else "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """The unmodified reference as PyTorch eager on one B200 -- same workload, same batch.  One JSON line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T = shape_of(args, 1)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    sampler = ClockSampler(dev.index)
    sampler.start()
    r = reference_steps(args.workload, B, T, args.steps, args.warmup, "cuda", autocast=args.ref_autocast, gl_iters=args.gl_iters)
    clocks = sampler.summary()
    if r is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref is missing (python baseline/install_reference.py)"}))
        return
    sps, sec, _ = r
    unit = "clips/s" if args.workload == "inference" else "samples/s"
    line = {"impl": "reference-gpu", "metric": METRIC[args.workload], "value": sps, "unit": unit, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "autocast bf16" if args.ref_autocast else "f32 (cuDNN TF32 convs, torch default)", "data": "synthetic",
            "config": dict(base_config(args, 1, B, T), how="unmodified reference modules, PyTorch eager, cudnn.benchmark=True "
                           "(train.py:53-54), stock step body incl. its host noise draw + H2D and its loss .item() sync"),
            "clocks": clocks, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(args, B, T, local):
    """Run `--impl reference-gpu` (fp32 and autocast-bf16) in child processes on the same GPU -> dict for the JSON line."""
    out = {}
    for name, extra in (("fp32", []), ("autocast_bf16", ["--ref-autocast"])):
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu", "--workload", args.workload, "--batch", str(B),
               "--frames", str(T), "--steps", str(args.eager_steps), "--warmup", "3", "--gl-iters", str(args.gl_iters)] + extra
        env = dict(os.environ, LOCAL_RANK=str(local), RANK="0", WORLD_SIZE="1")
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
            js = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            d = json.loads(js[-1]) if js else {"unavailable": (r.stderr or "no output")[-300:]}
        except Exception as e:   # noqa: BLE001
            d = {"unavailable": repr(e)[:300]}
        out[name] = {k: d.get(k) for k in ("value", "ms_per_step", "dtype", "unavailable", "peak_mem_gb") if d.get(k) is not None}
    return out


# -----------------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------------
def _dist_setup():
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
    return world, rank, local, dev, pg


def _traffic_record(key):
    """per-launch DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/roofline_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(key)
    except Exception:
        return None


def run_train(args):
    import vcagan_b200 as V
    from vcagan_b200.trainer import Trainer
    world, rank, local, dev, pg = _dist_setup()
    assert V.lib().query("vca_device_ok") == 1, "bench needs an sm_100 GPU (no fallback path exists)"
    lrs = args.workload == "lrs"
    B, T = shape_of(args, world)
    torch.manual_seed(1)
    V.manual_seed(1)
    tr = Trainer(precision=args.precision, dropout=True, device=dev, process_group=pg, lrs=lrs)   # broadcasts rank 0's weights
    vid_h, mel_h, spec_h, lens_h = synth(B, T, 100 + rank, lrs)
    vid_h, mel_h, spec_h = [t.pin_memory() for t in (vid_h, mel_h, spec_h)]
    vid, mel, spec = vid_h.to(dev), mel_h.to(dev), spec_h.to(dev)
    lens = lens_h.to(device=dev, dtype=torch.int32)
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
        tr.single_graph = tr.single_graph and not args.multi_graph
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

    # dominant kernel family: the tcgen05 implicit-GEMM conv (fwd + dgrad share one kernel); timed live per launch.
    # Every rank runs the instrumented step (it contains the gradient all-reduces); only rank 0 reports it
    # (serialised: the concurrent stream branches are switched off so that each launch is timed alone on its stream)
    tr.parallel_branches = False
    V.ops.cfg.param_grad_streams = ()
    prof = V.lib().profile_step(lambda: tr.step(vid, mel, spec, lens))
    barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if rank != 0:
        return
    pk = peaks()
    roof, exec_gf = None, None
    fam = {k: v for k, v in prof.items() if k.startswith("vca_conv_") and k.endswith(("_tc", "_tc_ws", "_tc_tm", "_tc_stats", "_tc_epi"))}
    flops = sum(v["flops"] for v in fam.values()); tms = sum(v["ms"] for v in fam.values())
    n_l = sum(v["n"] for v in fam.values())
    total_ms = sum(v["ms"] for v in prof.values())
    exec_gf = sum(v["flops"] for v in prof.values()) / B / 1e9      # FLOPs the library actually executed, per sample
    if tms > 0:
        ach = flops / (tms * 1e-3) / 1e12
        # the single geometry with the most time in the family
        best = max(((k, g) for k, v in fam.items() for g in v["top"]), key=lambda kg: kg[1][2], default=None)
        roof = {"bound": "tensor", "kernel": "conv_tc_fwd_kernel / conv_tc_ws_kernel / conv_tc_wgrad(_ws)_kernel (tcgen05 implicit GEMM family)",
                "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                "traffic": _traffic_record("train_dominant_conv_dram_bytes_per_launch"),
                "launches_per_step": n_l, "share_of_kernel_time": tms / total_ms, "peak_source": pk["src"] + " sustained",
                "family_flop_per_step": flops, "family_ms_per_step_serialised": tms,
                "top_geometry": None if best is None else {"entry": best[0], "geom": best[1][0], "launches": best[1][1],
                                                           "ms": best[1][2], "tflops": best[1][3]}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(sorted(prof.items(), key=lambda kv: -kv[1]["ms"])),
              open(os.path.join(ROOT, "gpurun_out", f"bench_kernel_breakdown_{args.workload}.json"), "w"), indent=1)
    sps = world * B / (ms * 1e-3)
    gf = GFLOP_PER_SAMPLE.get(T)
    eager = None
    if world == 1 and not args.no_gpu_eager:
        del tr
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(args, B, T, local)
        for k in ("fp32", "autocast_bf16"):
            if eager[k].get("value"):
                eager[k]["speedup_of_value"] = sps / eager[k]["value"]
                eager[k]["speedup_of_e2e"] = (world * B / (ms_e2e * 1e-3)) / eager[k]["value"]
        eager["how"] = ("unmodified reference modules (baseline/_ref), PyTorch eager on this GPU, cudnn.benchmark=True, stock "
                        f"train step body, same B={B} T={T}, 3 warm-up + {args.eager_steps} steps (bench.py --impl reference-gpu)")
    cpu = None
    if not args.no_cpu_baseline:
        c_sps, c_sec, c_kind, c_b = cpu_arm(args.workload, T, 1, 0)
        cpu = {"value": c_sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": c_kind,
               "sample": f"1 G+D step of B={c_b} (bounded sample), T={T}, fp32, "
                         + ("unmodified reference modules through the stock step body" if c_kind == "reference" else "oracle port")
                         + f" on torch CPU, all {os.cpu_count()} threads ({c_sec:.1f} s)"}
    if args.no_graph:
        launch = "eager"
    elif args.multi_graph and world == 1:
        launch = "3 CUDA graphs per step (D phase | G phase | G optimizer)"
    elif world == 1:
        launch = "1 CUDA graph per step"
    else:
        launch = tr_launch_desc()
    line = {
        "metric": METRIC[args.workload], "value": sps, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": dict(base_config(args, world, B, T), replicas_in_sync=in_sync, l2="256 MiB flush buffer written between timed steps",
                       launch=launch,
                       e2e_feed="per-step blocking copy" if args.no_graph else
                       "one timed region over all e2e steps; step i+1's H2D copy overlaps step i on a copy stream",
                       step_tensor_roofline_frac=(sps / world * gf * 1e9 / (pk["tf_sust"] * 1e12)) if gf else None,
                       algorithmic_gflop_per_sample=gf, executed_gflop_per_sample=exec_gf,
                       step_tensor_roofline_frac_executed=sps / world * exec_gf * 1e9 / (pk["tf_sust"] * 1e12)),
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": (vid_h.numel() + mel_h.numel() + spec_h.numel()) * 4, "d2h_bytes_per_step": 8},
        "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
    }
    print(json.dumps(line), flush=True)


def tr_launch_desc():
    return ("6 CUDA graphs per step (one per phase of Trainer._run_schedule), NCCL all-reduces issued between them on a comm stream: D "
            "grads underneath the Postnet forward + L1 terms, gen+post grads underneath the visual front-end backward, v_front grads "
            "underneath Adam on gen+post")


def run_inference(args):
    """BASELINE config[4]: test.py:126-143 on the device.  N > 1 = N independent replicas (no exchange step)."""
    import vcagan_b200 as V
    from vcagan_b200 import models as M, audio, infer
    world, rank, local, dev, pg = _dist_setup()
    B, T = shape_of(args, world)
    V.set_precision(args.precision)
    torch.manual_seed(1)
    V.manual_seed(1 + rank)
    vf, gen, post = M.Visual_front().to(dev).eval(), M.Decoder().to(dev).eval(), M.Postnet().to(dev).eval()
    g = torch.Generator().manual_seed(7 + rank)
    vid_h = torch.randn(B, 1, T, 112, 112, generator=g).pin_memory()
    vid = vid_h.to(dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_it = args.gl_iters

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        tot = 0.0
        for _ in range(n):
            l2_flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / n

    resident = lambda: infer.synthesize(vf, gen, post, vid, lens, n_iters=n_it, tta=True)                      # noqa: E731
    host = lambda: infer.synthesize(vf, gen, post, vid_h.to(dev, non_blocking=True), lens, n_iters=n_it, tta=True)["wav"].cpu()  # noqa: E731
    n0 = V.lib().launches
    for _ in range(max(args.warmup, 3)):
        resident()
    launches = (V.lib().launches - n0) // max(args.warmup, 3)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(resident, args.steps)
    barrier()
    clocks = sampler.summary()
    host()
    ms_e2e = timed(host, max(2, min(args.steps, 5)))
    # Griffin-Lim alone (the HBM-bound kernel the north star names), on a Postnet-shaped input
    Tp = 4 * T
    spec = torch.rand(B, 321, Tp, device=dev)
    audio.griffin_lim(spec, None, n_it)
    ms_gl = timed(lambda: audio.griffin_lim(spec, None, n_it), 3)
    ms_fwd = timed(lambda: infer.synthesize(vf, gen, post, vid, lens, n_iters=0, tta=True), 3)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    pk = peaks()
    alg_bytes = B * (n_it + 1) * (321 * Tp * 4 + 2 * 160 * (Tp - 1) * 4)   # mag read + signal read + write per clip-iteration
    ach = alg_bytes / (ms_gl * 1e-3) / 1e9
    gf = GFLOP_PER_CLIP_INFER.get(T)
    cpu = None
    if not args.no_cpu_baseline:
        c_sps, c_sec, c_kind, c_b = cpu_arm("inference", T, 1, 0, n_it)
        cpu = {"value": c_sps, "unit": "clips/s", "cores": os.cpu_count(), "kind": c_kind,
               "sample": f"1 pass over B={c_b} clips (bounded sample), T={T}: unmodified reference v_front + gen (x2, flip TTA) + post + "
                         f"its conv1d-based griffin_lim ({n_it} iterations) on torch CPU, all {os.cpu_count()} threads ({c_sec:.1f} s)"}
    eager = None
    if world == 1 and not args.no_gpu_eager:
        del vf, gen, post
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(args, B, T, local)
        for k in ("fp32", "autocast_bf16"):
            if eager[k].get("value"):
                eager[k]["speedup_of_value"] = (world * B / (ms * 1e-3)) / eager[k]["value"]
    line = {"metric": METRIC["inference"], "value": world * B / (ms * 1e-3), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 network / f32 Griffin-Lim" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": dict(base_config(args, world, B, T), l2="256 MiB flush buffer written between timed passes", launch="eager",
                           ms_forward_tta=ms_fwd, ms_griffin_lim=ms_gl,
                           forward_tensor_roofline_frac=(B / (ms_fwd * 1e-3) * gf * 1e9 / (pk["tf_sust"] * 1e12)) if gf else None),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": vid_h.numel() * 4,
                    "d2h_bytes_per_step": B * 160 * (Tp - 1) * 4},
            "roofline": {"bound": "hbm", "kernel": "Griffin-Lim (csrc/stft.cu)", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": ach / pk["hbm"], "traffic": _traffic_record("griffin_lim_dram_bytes_per_call"),
                         "algorithmic_bytes": alg_bytes, "peak_source": pk["src"]},
            "cpu_baseline": cpu, "gpu_eager_baseline": eager}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="train", choices=["train", "lrs", "inference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--global-batch", type=int, default=256, help="--scaling strong: the fixed global batch (BASELINE config[2])")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 32 train, 16 lrs, 64 inference)")
    ap.add_argument("--frames", type=int, default=0, help="frames per clip (default: 75 GRID, 250 lrs)")
    ap.add_argument("--lrs", action="store_true", help="same as --workload lrs")
    ap.add_argument("--gl-iters", type=int, default=60)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-GPU-eager denominator (N = 1 only)")
    ap.add_argument("--eager-steps", type=int, default=8)
    ap.add_argument("--ref-autocast", action="store_true", help="--impl reference-gpu under torch.autocast(bfloat16)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--multi-graph", action="store_true", help="one CUDA graph per phase with the all-reduces issued between them "
                                                               "(instead of one graph for the whole step, NCCL nodes included)")
    args = ap.parse_args()
    if args.lrs:
        args.workload = "lrs"
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    elif args.workload == "inference":
        run_inference(args)
    else:
        run_train(args)


if __name__ == "__main__":
    main()
