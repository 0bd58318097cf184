#!/bin/bash
# ROUND-1 script, kept for the record (bench.py now defaults to the 14B headline; use scripts/gpu_r02b.sh + gpu_round2.sh).
# One full GPU pass on 1 B200 (run under gpurun): parity tests, smoke, bench, kernel battery, ncu evidence.
# Usage: bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_$TAG.json
echo "== bench ti2v-5B (the reference's native model; not a BASELINE config)" ; timeout 600 python bench.py --config 5B --steps 3 --warmup 3 --skip-cpu 2>gpurun_out/bench_5B_$TAG.err | tee gpurun_out/bench_5B_$TAG.json | cut -c1-400
tail -2 gpurun_out/bench_5B_$TAG.err
echo "== bench tma-sweep" ; timeout 600 python bench.py --workload tma-sweep 2>gpurun_out/bench_sweep_$TAG.err | tee gpurun_out/bench_sweep_$TAG.json
tail -3 gpurun_out/bench_sweep_$TAG.err
echo "== kernel battery (prologue + cross)"
T=univid_b200/csrc/tests/uvb_test
for c in "prol 1 1950 12 1 0 0" "prol 2 300 12 1 1 0" "prol 1 1000 40 1 0 0" "prol 1 500 24 1 0 0" "prol 1 500 10 1 0 0" \
         "prol 1 32760 12 1 0 20" "prol 1 75600 40 1 0 10" "prol 1 27280 24 1 0 10" "fmha 1 32760 512 12 -1 0 10" "fmha 1 32760 32760 12 -1 0 5" \
         "gemm 1000 1536 1536 1 0" "gemm 32760 1536 1536 0 10" "gemm 32760 8960 1536 1 5" "gemm 32760 1536 8960 0 5" "gemm 75600 5120 5120 0 3"; do
  echo "-- $c"; timeout 120 $T $c 2>&1 | tail -3
done | tee gpurun_out/kernels_$TAG.log
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fmha_fwd_kernel|qk_norm_rope|head_scatter_kernel|block_glue|gemm_bf16" -c 900 \
    --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-denoise > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches_$TAG.csv)"
echo "== ncu --set full: self-attention kernel, prologue kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_fmha_$TAG \
    $T fmha 1 32760 32760 12 -1 0 1 > gpurun_out/ncu_fmha_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qk_norm_rope -s 1 -c 1 -f -o gpurun_out/prof_prol_$TAG \
    $T prol 1 32760 12 1 0 1 > gpurun_out/ncu_prol_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_xattn_$TAG \
    $T fmha 1 32760 512 12 -1 0 1 > gpurun_out/ncu_xattn_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 1 -c 1 -f -o gpurun_out/prof_gemm_$TAG \
    $T gemm 32760 8960 1536 1 1 > gpurun_out/ncu_gemm_$TAG.log 2>&1
echo "== ncu launch list of one full-size DiT block (every kernel, library ones included)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_block_$TAG.csv \
    python scripts/ncu_block.py > gpurun_out/ncu_block_$TAG.log 2>&1
echo "block launch list rows: $(wc -l < gpurun_out/launches_block_$TAG.csv)"
echo "== denoise step by category (CUDA events)"
python scripts/profile_denoise.py > gpurun_out/denoise_profile_$TAG.log 2>&1; cat gpurun_out/denoise_profile_$TAG.log
echo "== sampler step / batched CFG / per-token timesteps"
python scripts/bench_cfg_batch.py > gpurun_out/cfg_batch_$TAG.log 2>&1; tail -4 gpurun_out/cfg_batch_$TAG.log
# (round-1 in-step A/B of the FMA-pipe exp2 removed: the variant lives in the lab build only; see scripts/gpu_round2.sh for the round-2 A/Bs via bench.py --knob)
ls -la gpurun_out | tail -20
