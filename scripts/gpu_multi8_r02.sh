#!/bin/bash
# 8-GPU validation: the world-8 SP tests against the oracle, then bench at N=8 (headline 14B)
TAG=${TAG:-r02f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sp_gpu.py -x -q -m gpu -k "8-" > gpurun_out/pytest_sp_8gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_sp_8gpu_$TAG.log
tail -4 gpurun_out/pytest_sp_8gpu_$TAG.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/bench_8gpu_$TAG.json 2> gpurun_out/bench_8gpu_$TAG.err; echo "bench exit=$?"
tail -4 gpurun_out/bench_8gpu_$TAG.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_8gpu_$TAG.json 2>/dev/null; echo "ref exit=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_8gpu_$TAG.json", "gpurun_out/bench_ref_8gpu_$TAG.json"):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "denoise", d.get("denoise_step_ms"), "cores", (d.get("cpu_baseline") or {}).get("cores"))
    print(" parity", d.get("parity_check"))
    if d.get("kernel_split"): print(" split", {k: (round(v["avg_ms_per_layer"],4), round(v["min_over_ranks"],4), round(v["max_over_ranks"],4)) for k,v in d["kernel_split"]["segments"].items()})
    for k,v in (d.get("configs") or {}).items():
        if k=="tma_sweep": print(" sweep", {kk:(vv["kernel_ms_per_call"],vv["kernel_tflops"]) for kk,vv in v.items() if isinstance(vv,dict)})
PY
