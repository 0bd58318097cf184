#!/bin/bash
# A/B of fmha variants (UVB_FMHA_STEP x UVB_FMHA_POLY): correctness subset + timing. Output gpurun_out/fmha_ab.log
mkdir -p gpurun_out; T=univid_b200/csrc/tests/uvb_test; LOG=gpurun_out/fmha_ab.log; : > $LOG
VARIANTS=${VARIANTS:-128:0 128:4 128:2 64:0 64:4}
for V in $VARIANTS; do
  STEP=${V%%:*}; POLY=${V##*:}
  echo "##### UVB_FMHA_STEP=$STEP UVB_FMHA_POLY=$POLY" >> $LOG
  for c in "fmha 1 128 128 1 -1 0 0" "fmha 1 256 1024 2 -1 0 0" "fmha 1 300 77 2 -1 0 0" "fmha 2 1950 1950 3 1000 0 0" \
           "fmha 1 1950 512 12 -1 1 0" "fmha 1 4096 4096 4 -1 0 0" "fmha 1 32760 32760 12 -1 0 5" "fmha 1 32760 512 12 -1 0 10" "fmha 1 75600 75600 40 -1 0 2"; do
    echo "== $c" >> $LOG; UVB_FMHA_STEP=$STEP UVB_FMHA_POLY=$POLY timeout 120 $T $c >> $LOG 2>&1; echo "   exit=$?" >> $LOG
  done
done
grep -E "#####|PASS|FAIL|TIME|exit=[^0]" $LOG
