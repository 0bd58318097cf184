#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/xpair.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
T=univid_b200/csrc/tests/uvb_test
for k in 1 0; do
  echo "##### xattn_pair=$k" >> $LOG
  export UVB_KNOBS="xattn_pair=$k"
  run $T fmha 1 600 512 3 -1 0 0
  run $T fmha 1 600 512 3 -1 1 0
  run $T fmha 2 300 77 2 50 0 0
  run $T fmha 1 1950 1950 3 -1 0 0
  run $T fmha 1 32760 512 12 -1 0 10
  run $T fmha 1 75600 512 40 -1 0 5
  run $T fmha 1 8190 512 12 -1 0 10
  run $T fmha 1 32760 512 12 -1 1 10
done
unset UVB_KNOBS
for k in 2 1; do UVB_KNOBS="prologue_pair=$k" run $T prol 1 32760 12 1 0 20; UVB_KNOBS="prologue_pair=$k" run $T prol 1 75600 40 1 0 10; done
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py tests/test_animate_gpu.py -x -q -m gpu >> $LOG 2>&1
grep -v "exit=0" $LOG | cut -c1-250 | tail -60
