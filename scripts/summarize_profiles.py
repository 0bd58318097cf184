#!/usr/bin/env python
"""Turn the scratch artefacts of a GPU round (gpurun_out/) into the tracked evidence under profiles/<tag>/:
per-kernel aggregation of the ncu launch list, the key metrics of every `ncu --set full` capture, the bench
lines, and profiles/roofline_traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py).

    python scripts/summarize_profiles.py r01
"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
)


def raw_metrics(rep):
    """Metrics of every launch in a .ncu-rep, or in the `--page raw --csv` dump made of it on the GPU box."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, vals)} for vals in rows[2:]]


def launch_summary(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mi:
            continue
        v = float(r[mi].replace(",", ""))
        ms = v / 1e6 if r[ui].startswith("n") else (v / 1e3 if r[ui].startswith("u") else v)
        a = agg.setdefault((r[ki].split("(")[0], r[gi]), [0, 0.0])
        a[0] += 1
        a[1] += ms
    return agg


def to_bytes(val, unit):
    return float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles", tag)
    os.makedirs(dst, exist_ok=True)
    lines = [f"# profiles/{tag} -- evidence from GPU round `{tag}` (B200; one GPU unless the file name says otherwise)\n"]

    for f in sorted(glob.glob(os.path.join(src, f"bench*_{tag}.json")) + glob.glob(os.path.join(src, f"*_{tag}.log"))):
        if os.path.getsize(f) and "ncu_" not in os.path.basename(f):
            shutil.copy(f, dst)

    for lname, what in ((f"launches_{tag}.csv", "`bench.py --steps 2 --warmup 3 --skip-cpu --skip-denoise --no-sub-records` (headline config)"),
                        (f"launches_1p3B_{tag}.csv", "`bench.py --config 1.3B --steps 2 --warmup 3 --skip-cpu --skip-denoise --no-sub-records`")):
        lpath = os.path.join(src, lname)
        if not os.path.exists(lpath):
            continue
        shutil.copy(lpath, dst)
        agg = launch_summary(lpath)
        tot = sum(a[1] for a in agg.values())
        lines.append(f"\n## ncu launch list of {what} ({lname}; cold-cache, serialised)\n")
        lines.append("| kernel | grid | launches | total ms | share | avg us |\n|---|---|---:|---:|---:|---:|")
        for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"| `{name}` | {grid} | {n} | {t:.3f} | {100 * t / tot:.1f}% | {t / n * 1e3:.1f} |")

    bpath = os.path.join(src, f"launches_block_{tag}.csv")
    if os.path.exists(bpath):
        shutil.copy(bpath, dst)
        # keep the second pass only (after the second marker fill kernel)
        rows = list(csv.reader(open(bpath)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr, data = rows[hi], rows[hi + 1:]
        ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        # the block input x is created by a normal_ kernel etc.; passes are separated by 1-element fill kernels
        gi = hdr.index("Grid Size")
        marks = [i for i, r in enumerate(data) if len(r) > mi and "fill" in r[ki].lower() and r[gi].replace(" ", "") in ("(1,1,1)", "1,1,1")]
        start = marks[-1] + 1 if marks else 0
        agg = collections.OrderedDict()
        for r in data[start:]:
            if len(r) <= mi:
                continue
            v = float(r[mi].replace(",", ""))
            ms = v / 1e6 if r[ui].startswith("n") else (v / 1e3 if r[ui].startswith("u") else v)
            name = r[ki].split("(")[0]
            name = name if len(name) < 90 else name[:87] + "..."
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ms
        tot = sum(a[1] for a in agg.values()) or 1.0
        lines.append(f"\n## ncu launch list of ONE full-size DiT block forward, 1.3B config, 32 760 tokens (launches_block_{tag}.csv, second pass; every kernel, cold-cache, serialised)\n")
        lines.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"| `{name}` | {n} | {t:.3f} | {100 * t / tot:.1f}% |")
        lines.append(f"| total | {sum(a[0] for a in agg.values())} | {tot:.3f} | |")

    traffic = {}
    for rep in sorted(glob.glob(os.path.join(src, f"prof_*_{tag}.ncu-rep")) + glob.glob(os.path.join(src, f"prof_*_{tag}_rawpage.csv"))):
        name = os.path.basename(rep)
        name = name[:-len(".ncu-rep")] if name.endswith(".ncu-rep") else name[:-len("_rawpage.csv")]
        for m in raw_metrics(rep):
            kn = m.get("Kernel Name", ("?", ""))[0]
            lines.append(f"\n## `ncu --set full` {name} -- `{kn}` grid {m.get('Grid Size', ('?', ''))[0]} block {m.get('Block Size', ('?', ''))[0]}\n")
            lines.append("| metric | value | unit |\n|---|---:|---|")
            for k in KEYS:
                if k in m:
                    lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
            with open(os.path.join(dst, f"{name}_raw.csv"), "w") as fh:
                w = csv.writer(fh)
                for k, (v, u) in m.items():
                    w.writerow([k, v, u])
            if "dram__bytes_read.sum" in m:
                traffic[name] = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    if traffic:
        lines.append("\n## DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum)\n")
        for k, v in traffic.items():
            lines.append(f"- {k}: {v / 1e6:.1f} MB")
        tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        cur = json.load(open(tj)) if os.path.exists(tj) else {}
        for k, v in traffic.items():
            if k.startswith("prof_fmha14B_"):
                cur.setdefault("14B", {})["fmha_dram_bytes_per_launch"] = v
                cur["14B"]["source"] = f"profiles/{tag}/{k}_raw.csv"
            elif k.startswith("prof_prol14B_"):
                cur.setdefault("14B", {})["prologue_dram_bytes_per_launch"] = v
            elif k.startswith("prof_fmha_single_"):
                pass
            elif k.startswith("prof_fmha_"):
                cur.setdefault("1.3B", {})["fmha_dram_bytes_per_launch"] = v
                cur["1.3B"]["source"] = f"profiles/{tag}/{k}_raw.csv"
            elif k.startswith("prof_prol_"):
                cur.setdefault("1.3B", {})["prologue_dram_bytes_per_launch"] = v
            elif k.startswith("prof_gemm_"):
                cur.setdefault("1.3B", {})["gemm_ffn0_dram_bytes_per_launch"] = v
        json.dump(cur, open(tj, "w"), indent=1)
    open(os.path.join(dst, "SUMMARY.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
