#!/bin/bash
# final pass, driver-like: full GPU tests, smoke, bench and reference arm with the driver's --steps 20 --warmup 5
TAG=${TAG:-r02i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke_$TAG.log
tail -2 gpurun_out/smoke_$TAG.log
( time timeout 870 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit=$?"
tail -4 gpurun_out/bench_$TAG.err
( time timeout 870 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit=$?"
tail -4 gpurun_out/bench_ref_$TAG.err
python - <<PY
import json
for f in ("gpurun_out/bench_$TAG.json", "gpurun_out/bench_ref_$TAG.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "denoise", d.get("denoise_step_ms"), "steps", d["steps"], d["warmup"])
    if d.get("roofline"): print("  fmha", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), "prol", round(d["roofline_prologue"]["frac"],3), d["roofline_prologue"]["back_to_back"]["frac"], "gemm", round(d["roofline_gemm"]["frac"],3))
    s=(d.get("configs") or {}).get("1.3B")
    if s: print("  1.3B", round(s["value"],1), round(s["e2e"]["value"],1), round(s["denoise_step"]["ms"],1), s["denoise_step"]["launches_per_step"], round(s["roofline"]["achieved"],1), round(s["roofline_prologue"]["frac"],3))
    print("  clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
PY
