#!/bin/bash
# The TMA staging experiment for pass B (NKA_PASS_B_TMA=1) against the LDG.E.128 path: timing
# (events) and one ncu full capture of each kernel (DRAM throughput %, bytes).
tag=${1:-r2}
out=gpurun_out/tma_ab_$tag.jsonl
: > $out
for m in 10 5 20; do
  for tma in 0 1 0 1; do
    NKA_PASS_B_TMA=$tma TUNE_N=$((1<<28)) TUNE_M=$m TUNE_STEPS=40 TUNE_TAG="m$m tma=$tma" timeout 200 python tools/tune.py >> $out 2>> gpurun_out/tma_ab_$tag.err
  done
done
python - <<PY
import json
for ln in open("$out"):
    d = json.loads(ln); print("%-12s update %.4f ms  A %.3f  B %.4f ms (%.0f GB/s actual)" % (d["tag"], d["ms_update"], d["ms_a"], d["ms_b"], d["tbs_b_actual"]))
PY
for tma in 0 1; do
  NKA_PASS_B_TMA=$tma TUNE_N=$((1<<28)) TUNE_M=10 TUNE_STEPS=4 timeout 600 ncu --set full --clock-control none --import-source on \
     -k regex:nka_pass_b -s 16 -c 2 -f -o gpurun_out/prof_passb_tma${tma}_$tag python tools/tune.py > gpurun_out/ncu_passb_tma${tma}_$tag.log 2>&1
done
ls -la gpurun_out/prof_passb_tma*_$tag.ncu-rep
