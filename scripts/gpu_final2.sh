#!/bin/bash
# last pass of the round: sanitizer on the short-key pair kernel (changed last), then the driver-like pass
TAG=${TAG:-r02k}
mkdir -p gpurun_out
LOG=gpurun_out/sanitizer_xattn_$TAG.txt
: > $LOG
T=univid_b200/csrc/tests/uvb_test
CS=/usr/local/cuda/bin/compute-sanitizer
run() { echo "== $*" >> $LOG; timeout 300 "$@" 2>&1 | grep -v "^$" | tail -8 >> $LOG; echo "   exit=${PIPESTATUS[0]}" >> $LOG; }
for tool in memcheck racecheck synccheck; do
  echo "########## $tool" >> $LOG
  run $CS --tool $tool $T fmha 1 1100 512 3 -1 0 0
  run $CS --tool $tool $T fmha 1 600 512 3 -1 1 0
  run $CS --tool $tool $T fmha 2 1300 300 3 200 0 0
  run $CS --tool $tool $T fmha 2 300 77 2 50 0 0
done
grep "ERROR SUMMARY\|RACECHECK SUMMARY\|exit=" $LOG | sort | uniq -c | tail -12
bash scripts/gpu_final.sh
