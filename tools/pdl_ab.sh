#!/bin/bash
# A/B of programmatic dependent launch on the update chain (run on the GPU box): whole-update time
# without per-kernel events, PDL on vs off, at the sizes where launch boundaries matter.
out=${1:-gpurun_out/pdl_ab.jsonl}
: > $out
for cfg in "24 5" "25 10" "28 10"; do
  set -- $cfg
  for pdl in 1 0 1 0; do
    NKA_PDL=$pdl TUNE_SPANS=0 TUNE_N=$((1<<$1)) TUNE_M=$2 TUNE_STEPS=200 TUNE_TAG="n2^$1 m$2 pdl=$pdl" \
      timeout 200 python tools/tune.py >> $out 2>> ${out%.jsonl}.err
  done
done
python - <<PY
import json
for ln in open("$out"):
    d = json.loads(ln); print("%-22s update %.4f ms  (%.1f upd/s, frac %.3f)" % (d["tag"], d["ms_update"], d["updates_per_s"], d["frac_roofline"]))
PY
