#!/bin/bash
# Run-to-run spread of the n = 2^28 update (fresh process = fresh allocations each time), with the
# clocks sampled alongside: is a slow run a clock state or an allocation placement?
out=${1:-gpurun_out/alloc_repeat.jsonl}
: > $out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader -lms 250 > ${out%.jsonl}_smi.csv &
SMI=$!
for i in 1 2 3 4 5 6; do
  NKA_PDL=${NKA_PDL:-0} TUNE_SPANS=1 TUNE_N=$((1<<28)) TUNE_M=10 TUNE_STEPS=60 TUNE_TAG="run$i" timeout 200 python tools/tune.py >> $out 2>> ${out%.jsonl}.err
  date +%T.%N >> ${out%.jsonl}_marks.txt
done
kill $SMI
python - <<PY
import json
for ln in open("$out"):
    d = json.loads(ln); print("%-6s update %.4f ms  A %.3f B %.3f  frac %.3f" % (d["tag"], d["ms_update"], d["ms_a"], d["ms_b"], d["frac_roofline"]))
PY
