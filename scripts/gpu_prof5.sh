#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof5.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
T=univid_b200/csrc/tests/uvb_test
for k in "fmha_pair=1" "fmha_pair=1,fmha_poly=4" "fmha_pair=1,fmha_poly=3" "fmha_pair=1,fmha_poly=2" "fmha_pair=0"; do
  echo "##### $k" >> $LOG
  export UVB_KNOBS="$k"
  run $T fmha 1 4000 2500 2 -1 0 0
  run $T fmha 2 513 2200 3 100 0 0
  run $T fmha 1 32760 32760 12 -1 0 10
  run $T fmha 1 32760 32760 3 -1 0 10
  run $T fmha 1 75600 75600 5 -1 0 3
done
cat $LOG | cut -c1-220
