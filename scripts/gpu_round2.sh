#!/bin/bash
# Round-2 evidence pass on 1 B200 (run under gpurun): kernel battery, ncu launch lists + --set full captures, in-step A/Bs.
# Usage: bash scripts/gpu_round2.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=univid_b200/csrc/tests/uvb_test
echo "== kernel battery"
for c in "prol 1 32760 12 1 0 20" "prol 1 75600 40 1 0 10" "prol 1 27280 24 1 0 10" "fmha 1 32760 512 12 -1 0 10" "fmha 1 75600 512 40 -1 0 5" \
         "fmha 1 32760 32760 12 -1 0 5" "fmha 1 75600 75600 5 -1 0 3" \
         "gemm 32760 1536 1536 0 10" "gemm 32760 8960 1536 1 5" "gemm 75600 5120 5120 0 3"; do
  echo "-- $c"; timeout 120 $T $c 2>&1 | tail -2
done | tee gpurun_out/kernels_$TAG.log
echo "== in-step A/B: CTA-pair vs single-CTA attention, streaming vs token-pair prologue (1.3B config, 5 steps)"
for k in "fmha_pair=1" "fmha_pair=0" "prologue_pair=2" "prologue_pair=1" "fmha_pair=1" "fmha_pair=0"; do
  timeout 300 python bench.py --config 1.3B --steps 5 --warmup 3 --skip-cpu --skip-denoise --no-sub-records --knob $k 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$k value', round(d['value'],1), 'fmha in-step', round(d['roofline']['achieved'],1), 'prologue GB/s', round(d['roofline_prologue']['achieved']), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w_max'))"
done | tee gpurun_out/instep_ab_$TAG.log
echo "== ncu launch list of the bench command (headline 14B; the attention-stack step only)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fmha_fwd_kernel|qk_norm_rope|head_scatter_kernel|block_glue|gemm_bf16" -c 1000 \
    --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-denoise --no-sub-records > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches_$TAG.csv)"
echo "== ncu launch list, 1.3B config"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fmha_fwd_kernel|qk_norm_rope|head_scatter_kernel|block_glue|gemm_bf16" -c 900 \
    --csv --log-file gpurun_out/launches_1p3B_$TAG.csv python bench.py --config 1.3B --steps 2 --warmup 3 --skip-cpu --skip-denoise --no-sub-records > gpurun_out/ncu_bench_1p3B_$TAG.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches_1p3B_$TAG.csv)"
echo "== ncu --set full"
cap() {  # name kernel-regex args...
  local name=$1 rx=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o gpurun_out/prof_${name}_$TAG $T "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_${name}_$TAG.log
  # gpurun copies back at most 64 MiB: keep the raw metric page (what the summaries are made of), drop the 15 MB report
  ncu -i gpurun_out/prof_${name}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${name}_${TAG}_rawpage.csv 2>/dev/null
  rm -f gpurun_out/prof_${name}_$TAG.ncu-rep
}
cap fmha fmha_fwd_kernel fmha 1 32760 32760 12 -1 0 1
cap fmha14B fmha_fwd_kernel fmha 1 75600 75600 40 -1 0 1
cap prol qk_norm_rope prol 1 32760 12 1 0 1
cap prol14B qk_norm_rope prol 1 75600 40 1 0 1
cap xattn fmha_fwd_kernel fmha 1 32760 512 12 -1 0 1
cap gemm gemm_bf16_kernel gemm 32760 8960 1536 1 1
UVB_KNOBS="fmha_pair=0" cap fmha_single fmha_fwd_kernel fmha 1 32760 32760 12 -1 0 1
echo "== ncu launch list of one full-size DiT block"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_block_$TAG.csv \
    python scripts/ncu_block.py > gpurun_out/ncu_block_$TAG.log 2>&1
echo "block launch list rows: $(wc -l < gpurun_out/launches_block_$TAG.csv)"
echo "== denoise step by category (CUDA events)"
python scripts/profile_denoise.py > gpurun_out/denoise_profile_$TAG.log 2>&1; cat gpurun_out/denoise_profile_$TAG.log | tail -30
ls -la gpurun_out | tail -30
