#!/bin/bash
# Correctness + timing battery for the persistent (stream-K) attention kernel. Output gpurun_out/fmha_persist.log
mkdir -p gpurun_out; T=univid_b200/csrc/tests/uvb_test; LOG=gpurun_out/fmha_persist.log; : > $LOG
run() { echo "== $ENVV $*" >> $LOG; timeout 120 env $ENVV $T "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
for ENVV in "UVB_X=0" "UVB_TEST_NOWS=1"; do
  run fmha 1 128 128 1 -1 0 0
  run fmha 1 256 1024 2 -1 0 0
  run fmha 1 300 77 2 -1 0 0
  run fmha 2 1950 1950 3 -1 0 0
  run fmha 2 1950 1950 3 1000 0 0
  run fmha 3 700 1300 2 0 0 0
  run fmha 1 1950 512 12 -1 1 0
  run fmha 1 1950 1950 12 -1 0 3
  run fmha 1 4096 4096 4 -1 0 0
  run fmha 1 8190 8190 3 -1 0 3
  run fmha 1 40000 640 5 -1 1 3
done
ENVV="UVB_X=0"
run fmha 1 16380 16380 6 -1 0 0
run fmha 1 32760 512 12 -1 0 10
for ENVV in "UVB_X=0" "UVB_TEST_NOWS=1"; do
  run fmha 1 32760 32760 12 -1 0 5
  run fmha 1 32760 32760 6 -1 0 5
  run fmha 1 32760 32760 3 -1 0 5
done
ENVV="UVB_X=0"
run fmha 1 75600 75600 5 -1 0 3
run fmha 1 75600 75600 40 -1 0 2
grep -E "^==|PASS|FAIL|TIME|exit=[^0]" $LOG
