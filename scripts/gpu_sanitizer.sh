#!/bin/bash
# compute-sanitizer on the protocols that matter: the attention kernel's mbarrier / TMEM protocol
# (CTA-pair and single-CTA variants, split and unsplit schedules, ragged tails), the streaming prologue's bulk-copy
# ring, the GEMM, and the local-peer Ulysses exchange test.  Output: gpurun_out/sanitizer_$TAG.txt
TAG=${TAG:-r02}
mkdir -p gpurun_out
LOG=gpurun_out/sanitizer_$TAG.txt
: > $LOG
T=univid_b200/csrc/tests/uvb_test
CS=/usr/local/cuda/bin/compute-sanitizer
run() { echo "== $*" >> $LOG; timeout 600 "$@" 2>&1 | grep -v "^$" | tail -${TAILN:-12} >> $LOG; echo "   exit=${PIPESTATUS[0]}" >> $LOG; }
for tool in memcheck racecheck synccheck; do
  echo "########## $tool" >> $LOG
  # CTA-pair attention kernel: stream-K split (workspace) and unsplit, ragged last unit / key tile, k_lens
  run $CS --tool $tool $T fmha 1 700 2304 2 -1 0 0
  run $CS --tool $tool $T fmha 2 513 2200 3 100 0 0
  UVB_TEST_NOWS=1 run $CS --tool $tool $T fmha 1 700 2304 2 -1 0 0
  # single-CTA long-key kernel (knob) and the short-key (cross-attention) kernel with key modifiers
  UVB_KNOBS="fmha_pair=0" run $CS --tool $tool $T fmha 1 700 2304 2 -1 0 0
  run $CS --tool $tool $T fmha 1 600 512 3 -1 1 0
  run $CS --tool $tool $T fmha 2 300 77 2 50 0 0
  # streaming prologue (bulk-copy ring), all three widths, ragged last stage
  run $CS --tool $tool $T prol 1 1000 12 1 0 0
  run $CS --tool $tool $T prol 1 333 24 1 0 0
  run $CS --tool $tool $T prol 2 77 40 1 0 0
  # GEMM (CTA pairs)
  run $CS --tool $tool $T gemm 520 2296 200 0 0
done
echo "########## memcheck: local-peer Ulysses exchange (prologue/scatter peer stores, flag kernels, peer TMA stores)" >> $LOG
TAILN=6 run $CS --tool memcheck python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "sp_kernels_with_local_peers"
grep -c "ERROR SUMMARY: 0 errors" $LOG >> $LOG
grep "ERROR SUMMARY\|RACECHECK SUMMARY\|exit=" $LOG | sort | uniq -c | tail -20
