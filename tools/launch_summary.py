#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.   python tools/launch_summary.py in.csv [note]"""
import collections, csv, re, sys

def short(name):
    if "at::" in name:
        m = re.search(r"(FillFunctor|CUDAFunctor_add|CUDAFunctorOnSelf_add|direct_copy|CatArray|MulFunctor|reduce_kernel)", name)
        return "torch:" + (m.group(1) if m else "other")
    m = re.search(r"gemm_core<[^,>]*::(\w+)", name)
    if m:
        return "gemm_core<" + m.group(1) + ">"
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:50]

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")) / 1e6
    k = short(row["Kernel Name"]); agg[k][0] += 1; agg[k][1] += v; tot += v
print(f"# {' '.join(sys.argv[2:])}")
print(f"# total kernel time in window: {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches")
print(f"# {'ms':>8} {'share':>6} {'launches':>8} {'avg us':>8}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.3f} {100 * v[1] / tot:5.1f}% {v[0]:8d} {1e3 * v[1] / v[0]:8.1f}  {k}")
