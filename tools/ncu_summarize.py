#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the metrics DESIGN.md / bench.py quote, one row per captured launch:
    ncu -i gpurun_out/r02_kernels.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_summarize.py /tmp/raw.csv <tag>
writes profiles/ncu_<tag>_summary.csv and merges per-kernel DRAM bytes per launch into profiles/roofline_traffic.json."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "utchmma_bf16_pct_elapsed"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct_elapsed"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_mem_pct_elapsed"),
    ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sectors.sum", "l2_sectors"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("smsp__cycles_active.avg", "cycles_active"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
out, traffic = [], {}
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    kn = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").strip()
    rec = {"kernel": kn, "grid": r[col["Grid Size"]] if "Grid Size" in col else ""}
    for m, short in KEEP:
        cands = [n for n in names if n == m]
        if cands:
            c = col[cands[0]]
            rec[short] = r[c] + (" " + units[c] if units[c] and short in ("duration", "dram_read", "dram_write", "l2_bytes", "smem_dyn") else "")
    out.append(rec)


def to_bytes(s):
    v, u = (s.split() + [""])[:2]
    return float(v.replace(",", "")) * {"": 1, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


keys = ["kernel"] + [s for _, s in KEEP]
path = os.path.join(ROOT, "profiles", f"ncu_{tag}_summary.csv")
with open(path, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=keys, extrasaction="ignore")
    w.writeheader()
    for i, rec in enumerate(out):
        w.writerow(rec)
        if "dram_read" in rec and "dram_write" in rec:
            traffic[f"{i}:{rec['kernel']}"] = {"dram_bytes_per_launch": to_bytes(rec["dram_read"]) + to_bytes(rec["dram_write"]), "duration": rec.get("duration")}
tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
cur = json.load(open(tp)) if os.path.exists(tp) else {}
cur.setdefault("captures", {})[tag] = traffic
json.dump(cur, open(tp, "w"), indent=1)
print(open(path).read())
