#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prol.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
T=univid_b200/csrc/tests/uvb_test
for k in 2 1 0; do
  echo "##### prologue_pair=$k" >> $LOG
  export UVB_KNOBS="prologue_pair=$k"
  run $T prol 1 32760 12 1 0 20
  run $T prol 1 27280 24 1 0 20
  run $T prol 1 75600 40 1 0 10
  run $T prol 2 1000 12 1 0 0
done
unset UVB_KNOBS
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "prologue or qk_norm" >> $LOG 2>&1
cat $LOG | cut -c1-200
