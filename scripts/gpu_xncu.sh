#!/bin/bash
# one `ncu --set full` capture of the cross-attention pair kernel (cfg #2 shape) on the current code
TAG=${TAG:-r02k}
mkdir -p gpurun_out
T=univid_b200/csrc/tests/uvb_test
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_xattn_$TAG $T fmha 1 32760 512 12 -1 0 1 > gpurun_out/ncu_xattn_$TAG.log 2>&1
tail -2 gpurun_out/ncu_xattn_$TAG.log
ncu -i gpurun_out/prof_xattn_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_xattn_${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_xattn_$TAG.ncu-rep
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/prof_xattn_${TAG}_raw.csv")))
h=rows[0]; v=rows[-1]
for k in ("gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__cycles_active.avg"):
    if k in h: print(k, v[h.index(k)], rows[1][h.index(k)])
PY
