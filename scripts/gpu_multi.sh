#!/bin/bash
# Multi-GPU pass (run under gpurun --gpus N): Ulysses parity test + bench at N ranks.
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo_${N}gpu_$TAG.log
echo "== pytest sp"; timeout 900 python -m pytest tests/test_sp_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_sp_${N}gpu_$TAG.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  echo "== bench --gpus $n"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --skip-cpu --skip-denoise 2>gpurun_out/bench_${n}gpu_$TAG.err | tee gpurun_out/bench_${n}gpu_$TAG.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $n --steps 3 --warmup 3 --skip-cpu --skip-denoise 2>gpurun_out/bench_${n}gpu_$TAG.err | tee gpurun_out/bench_${n}gpu_$TAG.json
  fi
  tail -3 gpurun_out/bench_${n}gpu_$TAG.err
done
