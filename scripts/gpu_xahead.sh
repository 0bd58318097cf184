#!/bin/bash
# cross-unit early score issue in the short-key (cross-attention) pair kernel: correctness + timing + wait-cycle profile
mkdir -p gpurun_out
LOG=gpurun_out/xahead.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
P=univid_b200/csrc/tests/prof/uvb_test
T=univid_b200/csrc/tests/uvb_test
for k in 1 0; do
  echo "##### xattn_pair=$k" >> $LOG
  export UVB_KNOBS="xattn_pair=$k"
  run $T fmha 1 600 512 3 -1 0 0
  run $T fmha 1 600 512 3 -1 1 0
  run $T fmha 2 300 77 2 50 0 0
  run $T fmha 2 1300 300 3 200 0 0
  run $T fmha 1 1950 1950 3 -1 0 0
  run $T fmha 1 32760 257 24 -1 0 5
  run $T fmha 1 32760 512 12 -1 0 10
  run $T fmha 1 75600 512 40 -1 0 5
  run $T fmha 1 8190 512 12 -1 0 10
  run $T fmha 1 32760 512 12 -1 1 10
done
unset UVB_KNOBS
run $P fmha 1 32760 512 12 -1 0 5
run $P fmha 1 75600 512 40 -1 0 5
UVB_TEST_TIMELINE=1 run $T fmha 1 32760 512 12 -1 0 0
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_animate_gpu.py tests/test_full_size_gpu.py -x -q -m gpu >> $LOG 2>&1
echo "pytest exit=$?" >> $LOG
grep -v "^TL \|exit=0" $LOG | cut -c1-330 | tail -70
grep "^TL " $LOG | head -6
