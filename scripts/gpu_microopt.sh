#!/bin/bash
# cycle-level A/B of a softmax micro-optimisation: wait-cycle profile (cycles per step are clock-independent) + timing
mkdir -p gpurun_out
LOG=gpurun_out/microopt.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
P=univid_b200/csrc/tests/prof/uvb_test
T=univid_b200/csrc/tests/uvb_test
run $T fmha 1 4000 2500 2 -1 0 0
run $T fmha 2 513 2200 3 100 0 0
run $T fmha 1 600 512 3 -1 1 0
run $T fmha 3 129 513 1 1 0 0
for i in 1 2; do
run $P fmha 1 32760 32760 12 -1 0 5
UVB_KNOBS="fmha_pair=0" run $P fmha 1 32760 32760 12 -1 0 5
run $P fmha 1 32760 512 12 -1 0 5
done
run $T fmha 1 32760 32760 12 -1 0 10
run $T fmha 1 75600 75600 5 -1 0 3
run $T fmha 1 32760 512 12 -1 0 10
grep "PASS\|FAIL\|per step\|TIME" $LOG | cut -c1-200
