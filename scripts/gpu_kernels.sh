#!/bin/bash
# Stand-alone kernel battery on one B200 (run under gpurun). Each case is its own process with a
# timeout so a trapped/hung kernel cannot take the rest down. Output: gpurun_out/kernels.log
mkdir -p gpurun_out
T=univid_b200/csrc/tests/uvb_test
LOG=gpurun_out/kernels.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 120 env $ENVV $T "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
ENVV=""
run prol 1 1950 12 1 0 0
run prol 2 300 12 1 1 0
run prol 1 1000 40 1 0 0
run prol 1 500 24 0 0 0
run prol 1 500 10 1 0 0
run fmha 1 128 128 1 -1 0 0
ENVV="UVB_INMODE=1" run fmha 1 128 128 1 -1 0 0
ENVV="UVB_INMODE=2" run fmha 1 128 128 1 -1 0 0
ENVV="UVB_INMODE=3" run fmha 1 128 128 1 -1 0 0
ENVV=""
run fmha 1 256 256 1 -1 0 0
run fmha 1 256 1024 2 -1 0 0
run fmha 1 300 77 2 -1 0 0
run fmha 2 1950 1950 3 -1 0 0
run fmha 2 1950 1950 3 1000 0 0
run fmha 1 1950 512 12 -1 1 0
run fmha 1 4096 4096 4 -1 0 0
run prol 1 32760 12 1 0 20
run prol 1 75600 40 1 0 10
run fmha 1 32760 32760 12 -1 0 5
run fmha 1 32760 512 12 -1 1 10
run fmha 1 75600 75600 40 -1 0 2
cat $LOG
