#!/bin/bash
# round-2 evidence pass on one GPU: full GPU test suite, smoke, bench (headline 14B + sub-records), reference arm
mkdir -p gpurun_out
TAG=${TAG:-r02b}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke_$TAG.log
tail -3 gpurun_out/smoke_$TAG.log
( time timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit=$?"
tail -3 gpurun_out/bench_$TAG.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit=$?"
tail -3 gpurun_out/bench_ref_$TAG.err
python - <<'PY'
import json,glob,os
tag=os.environ.get("TAG","r02b")
for f in (f"gpurun_out/bench_{tag}.json", f"gpurun_out/bench_ref_{tag}.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "denoise", d.get("denoise_step_ms"))
    if d.get("roofline"): print("  roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "prol", (d.get("roofline_prologue") or {}).get("frac"))
    if d.get("kernel_split"): print("  split", {k: round(v["avg_ms_per_layer"],4) for k,v in d["kernel_split"]["segments"].items()})
    for k,v in (d.get("configs") or {}).items():
        if k=="tma_sweep": print("  sweep", {kk:(vv["kernel_ms_per_call"],vv["kernel_tflops"]) for kk,vv in v.items() if isinstance(vv,dict)})
        else: print("  sub",k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), (v.get("denoise_step") or {}).get("ms"), (v.get("roofline") or {}).get("achieved"), (v.get("roofline_prologue") or {}).get("frac"))
    print("  clocks", d.get("clocks"), "cpu", d.get("cpu_baseline"))
PY
