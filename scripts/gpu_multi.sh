#!/bin/bash
# Multi-GPU pass (run under gpurun --gpus N): Ulysses parity test + bench at 1..N ranks, both exchange transports.
# Usage: bash scripts/gpu_multi.sh N [tag] [configs...]
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo_${N}gpu_$TAG.log
echo "== pytest sp"; timeout 900 python -m pytest tests/test_sp_gpu.py -x -q -rs 2>&1 | tail -15 | tee gpurun_out/pytest_sp_${N}gpu_$TAG.log
bench() {  # n transport config
  local n=$1 tr=$2 cfg=$3 out=gpurun_out/bench_${3}_${1}gpu_${2}_$TAG
  echo "== bench --gpus $n ($tr, $cfg)"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --config $cfg --steps 3 --warmup 3 --skip-cpu --skip-denoise 2>$out.err | tee $out.json
  else
    UVB_SP_P2P=$([ $tr = p2p ] && echo 1 || echo 0) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
      --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --config $cfg --steps 3 --warmup 3 --skip-cpu --skip-denoise \
      2>$out.err | tee $out.json
  fi
  tail -3 $out.err
}
for cfg in ${@:3}; do
  for n in 1 2 4 8; do
    [ $n -le $N ] || continue
    if [ $cfg = 1.3B ] && [ $n -eq 8 ]; then continue; fi
    bench $n p2p $cfg
    [ $n -gt 1 ] && [ "${NCCL_ARM:-1}" = 1 ] && bench $n nccl $cfg
  done
done
