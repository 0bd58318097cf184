#!/bin/bash
# wait-cycle profile with the unit set-up / epilogue split (UVB_FMHA_PROFILE build): cross-attention shapes and, for
# comparison, the long-key self-attention kernel
mkdir -p gpurun_out
LOG=gpurun_out/xprof2_${TAG:-r02n}.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
P=univid_b200/csrc/tests/prof/uvb_test
run $P fmha 1 32760 512 12 -1 0 5
run $P fmha 1 75600 512 40 -1 0 5
run $P fmha 1 32760 32760 12 -1 0 3
UVB_KNOBS="xattn_pair=0" run $P fmha 1 32760 512 12 -1 0 5
cat $LOG | cut -c1-400
