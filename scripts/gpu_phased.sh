#!/bin/bash
# phased exchange validation on N GPUs: kernel tests (1 GPU part), SP tests vs oracle, bench with and without phases
N=${N:-2}
TAG=${TAG:-r02d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_linear_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_sp_gpu.py -x -q -m gpu > gpurun_out/pytest_sp_${N}gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_sp_${N}gpu_$TAG.log
tail -4 gpurun_out/pytest_sp_${N}gpu_$TAG.log
for ph in 1 0; do
  ( UVB_SP_PHASED=$ph timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ph bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --skip-cpu ) > gpurun_out/bench_${N}gpu_phased${ph}_$TAG.json 2> gpurun_out/bench_${N}gpu_phased${ph}_$TAG.err; echo "bench phased=$ph exit=$?"
  tail -3 gpurun_out/bench_${N}gpu_phased${ph}_$TAG.err
  python - <<PY
import json
f="gpurun_out/bench_${N}gpu_phased${ph}_$TAG.json"
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
except Exception as e:
    print("unparsable", e); raise SystemExit
print("phased=$ph value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "denoise", d.get("denoise_step_ms"), "parity", (d.get("parity_check") or {}).get("ok"))
if d.get("kernel_split"): print("split", {k: round(v["avg_ms_per_layer"],4) for k,v in d["kernel_split"]["segments"].items()})
for k,v in (d.get("configs") or {}).items():
    if k!="tma_sweep":
        print("sub",k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), (v.get("denoise_step") or {}).get("ms"), (v.get("parity_check") or {}).get("ok"))
        if v.get("kernel_split"): print("  split", {kk: round(vv["avg_ms_per_layer"],4) for kk,vv in v["kernel_split"]["segments"].items()})
PY
done
