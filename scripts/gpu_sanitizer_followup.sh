mkdir -p gpurun_out
T=univid_b200/csrc/tests/uvb_test
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "== synccheck xattn keymod"; timeout 300 $CS --tool synccheck $T fmha 1 600 512 3 -1 1 0 2>&1 | grep -v "Host Frame" | head -60
echo "== synccheck xattn plain short keys"; timeout 300 $CS --tool synccheck $T fmha 1 600 512 3 -1 0 0 2>&1 | grep -v "Host Frame" | head -30
echo "== racecheck pair"; timeout 300 $CS --tool racecheck $T fmha 1 700 2304 2 -1 0 0 2>&1 | tail -8
echo "== racecheck gemm pair"; timeout 300 $CS --tool racecheck $T gemm 520 2296 200 0 0 2>&1 | tail -8
} > gpurun_out/sanitizer_followup_r02.txt 2>&1
cat gpurun_out/sanitizer_followup_r02.txt | cut -c1-250
