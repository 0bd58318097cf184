#!/bin/bash
# A/B of fmha variants (UVB_FMHA_STEP) : correctness subset + timing. Output gpurun_out/fmha_ab.log
mkdir -p gpurun_out; T=univid_b200/csrc/tests/uvb_test; LOG=gpurun_out/fmha_ab.log; : > $LOG
for STEP in 128 64; do
  echo "##### UVB_FMHA_STEP=$STEP" >> $LOG
  for c in "fmha 1 128 128 1 -1 0 0" "fmha 1 256 1024 2 -1 0 0" "fmha 1 300 77 2 -1 0 0" "fmha 2 1950 1950 3 1000 0 0" "fmha 2 1950 1950 3 -1 0 0" \
           "fmha 1 1950 512 12 -1 1 0" "fmha 1 4096 4096 4 -1 0 0" "fmha 1 32760 32760 12 -1 0 5" "fmha 1 32760 512 12 -1 0 10" "fmha 1 75600 75600 40 -1 0 2"; do
    echo "== $c" >> $LOG; UVB_FMHA_STEP=$STEP timeout 120 $T $c >> $LOG 2>&1; echo "   exit=$?" >> $LOG
  done
done
grep -E "#####|PASS|FAIL|TIME|exit=[^0]" $LOG
