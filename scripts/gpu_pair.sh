#!/bin/bash
# CTA-pair attention kernel: correctness (stand-alone binary, naive fp32 kernel as the checker) + A/B timing against
# the single-CTA kernel.  Output: gpurun_out/pair.log
mkdir -p gpurun_out
LOG=gpurun_out/pair.log
: > $LOG
T=univid_b200/csrc/tests/uvb_test
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
# correctness: small shapes first (a protocol bug traps after a few seconds instead of hanging)
run $T fmha 1 512 2304 1 -1 0 0
run $T fmha 1 2100 2100 1 -1 0 0
run $T fmha 1 4000 2500 2 -1 0 0
run $T fmha 2 513 2200 3 100 0 0
run $T fmha 1 130 3000 2 -1 0 0
run $T fmha 1 9000 4000 20 -1 0 0
UVB_TEST_NOWS=1 run $T fmha 1 4000 2500 2 -1 0 0
# timing at the benchmark geometries: pair vs single
for k in 1 0; do
  echo "##### fmha_pair=$k" >> $LOG
  export UVB_KNOBS="fmha_pair=$k"
  run $T fmha 1 32760 32760 12 -1 0 10
  run $T fmha 1 32760 32760 6 -1 0 10
  run $T fmha 1 32760 32760 3 -1 0 10
  run $T fmha 1 75600 75600 5 -1 0 3
done
unset UVB_KNOBS
echo "##### pytest" >> $LOG
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_full_size_gpu.py -x -q -m gpu >> $LOG 2>&1
echo "   exit=$?" >> $LOG
tail -60 $LOG
