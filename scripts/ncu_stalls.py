#!/usr/bin/env python
"""Per-instruction stall summary of an ncu report (source page): python scripts/ncu_stalls.py <rep> [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
m = dict(zip(rows[0], rows[2]))
for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"):
    print(f"{k} = {m.get(k)}")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; agg = {}; recs = []
for r in data:
    try: n = int(r[ix["# Samples"]])
    except Exception: continue
    tot += n; recs.append((n, r))
    for c in stall_cols: agg[c] = agg.get(c, 0) + int(r[ix[c]] or 0)
print("total samples", tot)
print("  ".join(f"{c[6:]}={100*v/tot:.1f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
recs.sort(key=lambda t: -t[0])
for n, r in recs[:top]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{100*n/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:64]:64s} {st}")
