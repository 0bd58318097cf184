#!/bin/bash
# N-GPU validation: SP tests against the oracle, then bench at N (headline 14B + sub-records)
N=${N:-2}
TAG=${TAG:-r02c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sp_gpu.py -x -q -m gpu > gpurun_out/pytest_sp_${N}gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_sp_${N}gpu_$TAG.log
tail -4 gpurun_out/pytest_sp_${N}gpu_$TAG.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 ) > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench exit=$?"
tail -5 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
f="gpurun_out/bench_${N}gpu_$TAG.json"
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
except Exception as e:
    print("unparsable", e); raise SystemExit
print("value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "denoise", d.get("denoise_step_ms"))
print("parity", d.get("parity_check"))
if d.get("kernel_split"): print("split", {k: (round(v["avg_ms_per_layer"],4), round(v["min_over_ranks"],4), round(v["max_over_ranks"],4)) for k,v in d["kernel_split"]["segments"].items()})
for k,v in (d.get("configs") or {}).items():
    if k=="tma_sweep": print("sweep", {kk:(vv["kernel_ms_per_call"],vv["kernel_tflops"]) for kk,vv in v.items() if isinstance(vv,dict)})
    else:
        print("sub",k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), (v.get("denoise_step") or {}).get("ms"), v.get("parity_check"))
        if v.get("kernel_split"): print("  split", {kk: (round(vv["avg_ms_per_layer"],4), round(vv["min_over_ranks"],4), round(vv["max_over_ranks"],4)) for kk,vv in v["kernel_split"]["segments"].items()})
PY
