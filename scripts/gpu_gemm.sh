#!/bin/bash
# Stand-alone battery for uvb_linear_bf16 on one B200 (run under gpurun): single-CTA and CTA-pair kernels, both
# tile widths, each case its own process with a timeout. Output: gpurun_out/gemm.log
mkdir -p gpurun_out
T=univid_b200/csrc/tests/uvb_test
LOG=gpurun_out/gemm.log
: > $LOG
run() { echo "== [$ENVV] $*" >> $LOG; timeout 120 env $ENVV $T "$@" >> $LOG 2>&1; echo "   exit=$?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for c in 1 2; do
  for bn in 192 256; do
    ENVV="UVB_KNOBS=gemm_ctas=$c,gemm_bn=$bn"      # the TEST binary maps UVB_KNOBS onto uvb_set_knob
    run gemm 128 256 64 0 0
    run gemm 256 512 1536 0 0
    run gemm 1000 1536 1536 1 0
    run gemm 520 2296 200 1 0
    run gemm 32760 1536 1536 0 10
    run gemm 32760 1536 8960 0 5
  done
  ENVV="UVB_KNOBS=gemm_ctas=$c"
  run gemm 4096 4096 4096 0 5
  run gemm 32760 8960 1536 1 5
  run gemm 75600 5120 5120 0 3
  run gemm 75600 5120 13824 0 3
  run gemm 75600 13824 5120 1 3
done
ENVV="UVB_KNOBS="
run gemm 512 1536 1536 0 20
run gemm 512 5120 5120 0 20
run gemm 1 1536 1536 1 0
run gemm 130 200 72 1 0
ENVV="UVB_KNOBS=gemm_small=0"
run gemm 512 1536 1536 0 20
cat $LOG
