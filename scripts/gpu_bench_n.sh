#!/bin/bash
# bench at N GPUs only (final-code refresh of the scaling rows)
N=${N:-2}; TAG=${TAG:-r02h}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench exit=$?"
tail -4 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${N}gpu_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N", d["n_gpus"], "value", round(d["value"],1), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "denoise", round(d["denoise_step_ms"],1), "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_vs_fp32_rows"])
print(" split", {k: round(v["avg_ms_per_layer"],4) for k,v in d["kernel_split"]["segments"].items()})
s=d["configs"].get("1.3B")
if s: print(" 1.3B", round(s["value"],1), round(s["ms_per_step"],2), round(s["e2e"]["value"],1), round(s["denoise_step"]["ms"],1), {k: round(v["avg_ms_per_layer"],4) for k,v in s["kernel_split"]["segments"].items()})
print(" sweep", {k:(round(v["kernel_ms_per_call"],4)) for k,v in d["configs"]["tma_sweep"].items() if isinstance(v,dict)})
PY
