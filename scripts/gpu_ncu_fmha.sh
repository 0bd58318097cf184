#!/bin/bash
# ncu --set full capture of the self-attention kernel (stand-alone binary). Usage: gpu_ncu_fmha.sh <tag> [L] [N]
TAG=${1:-x}; L=${2:-32760}; N=${3:-12}
mkdir -p gpurun_out
T=univid_b200/csrc/tests/uvb_test
timeout 120 $T fmha 1 $L $L $N -1 0 5 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_fmha_$TAG \
    $T fmha 1 $L $L $N -1 0 1 > gpurun_out/ncu_fmha_$TAG.log 2>&1
tail -3 gpurun_out/ncu_fmha_$TAG.log
