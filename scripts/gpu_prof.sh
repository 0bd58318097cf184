#!/bin/bash
# barrier-wait cycle counters of the attention kernel (lab build with -DUVB_FMHA_PROFILE): pair vs single CTAs
mkdir -p gpurun_out
LOG=gpurun_out/prof.log
: > $LOG
T=univid_b200/csrc/tests/prof/uvb_test
for k in ${KNOB_LIST:-1 0}; do
  echo "##### fmha_pair=$k" >> $LOG
  for c in "1 32760 32760 12 -1 0 5" "1 32760 32760 3 -1 0 5"; do
    echo "== fmha $c" >> $LOG
    UVB_KNOBS="fmha_pair=$k" timeout 120 $T fmha $c >> $LOG 2>&1; echo "   exit=$?" >> $LOG
  done
done
cat $LOG
