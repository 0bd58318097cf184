#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof2.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
export UVB_KNOBS="fmha_pair=1"
run univid_b200/csrc/tests/prof/uvb_test fmha 1 32760 32760 12 -1 0 5
T=univid_b200/csrc/tests/uvb_test
run $T fmha 1 4000 2500 2 -1 0 0
run $T fmha 2 513 2200 3 100 0 0
run $T fmha 1 32760 32760 12 -1 0 10
run $T fmha 1 32760 32760 3 -1 0 10
run $T fmha 1 75600 75600 5 -1 0 3
cat $LOG
