#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/xprof.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
P=univid_b200/csrc/tests/prof/uvb_test
T=univid_b200/csrc/tests/uvb_test
run $P fmha 1 32760 512 12 -1 0 5
run $P fmha 1 75600 512 40 -1 0 5
UVB_TEST_TIMELINE=1 run $T fmha 1 32760 512 12 -1 0 0
for k in 2 1; do UVB_KNOBS="prologue_pair=$k" run $T prol 1 32760 12 1 0 20; UVB_KNOBS="prologue_pair=$k" run $T prol 1 75600 40 1 0 10; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "prologue or qk_norm" >> $LOG 2>&1
grep -v "^TL " $LOG | cut -c1-330
grep "^TL " $LOG | head -12
