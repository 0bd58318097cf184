#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof4.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
T=univid_b200/csrc/tests/uvb_test
for k in 2; do
  echo "##### fmha_pair=$k" >> $LOG
  export UVB_KNOBS="fmha_pair=$k"
  run $T fmha 1 512 2304 1 -1 0 0
  run $T fmha 1 2100 2100 1 -1 0 0
  run $T fmha 1 4000 2500 2 -1 0 0
  run $T fmha 2 513 2200 3 100 0 0
  run $T fmha 1 130 3000 2 -1 0 0
  run $T fmha 1 9000 4000 20 -1 0 0
  UVB_TEST_NOWS=1 run $T fmha 1 4000 2500 2 -1 0 0
  run univid_b200/csrc/tests/prof/uvb_test fmha 1 32760 32760 12 -1 0 5
  run $T fmha 1 32760 32760 12 -1 0 10
  run $T fmha 1 32760 32760 6 -1 0 10
  run $T fmha 1 32760 32760 3 -1 0 10
  run $T fmha 1 75600 75600 5 -1 0 3
done
cat $LOG | cut -c1-220
