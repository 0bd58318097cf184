#!/bin/bash
# 8-GPU validation pass (run under gpurun --gpus 8): Ulysses parity at 4 and 8 ranks, bench of the 14B config at
# 8 / 4 ranks and of the 1.3B config at 4 ranks.  Every command is bounded by its own timeout.
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo_8gpu_$TAG.log
echo "== pytest sp (4 ranks both transports, 8 ranks peer transport)"
timeout 600 python -m pytest tests/test_sp_gpu.py -x -q -rs -k "4 or 8-p2p" 2>&1 | tail -8 | tee gpurun_out/pytest_sp_8gpu_$TAG.log
bench() {  # n transport config extra-flags
  local n=$1 tr=$2 cfg=$3; shift 3
  local out=gpurun_out/bench_${cfg}_${n}gpu_${tr}_$TAG
  echo "== bench --gpus $n ($tr, $cfg) $*"
  UVB_SP_P2P=$([ $tr = p2p ] && echo 1 || echo 0) timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
    --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --config $cfg --steps 3 --warmup 3 --skip-cpu "$@" \
    2>$out.err | tee $out.json | cut -c1-400
  grep -E "Error|error|Traceback" $out.err | head -5
}
bench 8 p2p 14B
bench 8 nccl 14B --skip-denoise
bench 4 p2p 14B --skip-denoise
bench 4 p2p 1.3B
bench 4 nccl 1.3B --skip-denoise
