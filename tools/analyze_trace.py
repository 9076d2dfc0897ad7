#!/usr/bin/env python
"""Summarise gpurun_out/step_trace_<tag>.json (tools/step_trace.py): per kernel family the in-graph busy time, the
time during which it was the ONLY thing running, idle gaps, and concurrency over the step.
    python tools/analyze_trace.py gpurun_out/step_trace_r02.json"""
import json, re, sys
from collections import defaultdict

ks = json.load(open(sys.argv[1]))
span = max(ts + d for _, _, ts, d in ks)


def fam(n):
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"<.*", "", n)
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if n.startswith("torch:"):
        return n[:40]
    if n.startswith("at::native") or "at::" in n:
        return "torch:" + n.split("::")[-1][:30]
    return n[:40]


# sweep line over kernel start/end events
evs = []
for i, (n, s, ts, d) in enumerate(ks):
    evs.append((ts, 1, i)); evs.append((ts + d, 0, i))
evs.sort()
active = set(); last = 0.0
idle = 0.0; conc_time = defaultdict(float); alone = defaultdict(float); busy = defaultdict(float); cnt = defaultdict(int)
for t, kind, i in evs:
    dt = t - last
    if dt > 0:
        if not active:
            idle += dt
        conc_time[min(len(active), 8)] += dt
        if len(active) == 1:
            alone[fam(ks[next(iter(active))][0])] += dt
    last = t
    if kind == 1:
        active.add(i)
    else:
        active.discard(i)
for n, s, ts, d in ks:
    busy[fam(n)] += d; cnt[fam(n)] += 1
print(f"span {span / 1e3:.2f} ms; {len(ks)} kernels on {len(set(k[1] for k in ks))} streams; no kernel running: {idle / 1e3:.2f} ms")
print("time with k kernels in flight:", {k: round(v / 1e3, 2) for k, v in sorted(conc_time.items())})
print(f"{'family':42s} {'n':>5s} {'busy ms':>9s} {'alone ms':>9s}")
for f, b in sorted(busy.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{f:42s} {cnt[f]:5d} {b / 1e3:9.3f} {alone[f] / 1e3:9.3f}")
print("sum busy", round(sum(busy.values()) / 1e3, 2), "ms; sum alone", round(sum(alone.values()) / 1e3, 2), "ms")
# idle gaps > 20 us
gaps = []
active = 0; last_end = 0.0
cur_end = 0.0
for n, s, ts, d in ks:
    if ts > cur_end + 20 and cur_end > 0:
        gaps.append((round(cur_end / 1e3, 3), round((ts - cur_end), 1), n[:40]))
    cur_end = max(cur_end, ts + d)
print("idle gaps > 20 us (at ms, us, next kernel):", gaps[:30])
