#!/bin/bash
# Residual strip kernel: item mapping (absolute vs per-strip bands) and band height, 4096^2 and 8192^2 (GPU box).
tag=$1
out=gpurun_out/residual_bands_$tag.jsonl
: > $out
timeout 600 python -m pytest tests/test_gpu_example.py -m gpu -q -x -k "residual or golden or full_size" > gpurun_out/pytest_residual_$tag.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_residual_$tag.log)"
for abs in 0 1; do
  for band in 0 64 128 256; do
    for N in 4096 8192; do
      export NKA_RES_ABS_BANDS=$abs
      if [ $band = 0 ]; then unset NKA_RES_BAND; else export NKA_RES_BAND=$band; fi
      echo -n "{\"abs\": $abs, \"band\": $band, \"run\": " >> $out
      timeout 300 python tools/example_time.py $N 10 5 >> $out 2>> ${out%.jsonl}.err
      sed -i '$ s/$/}/' $out
    done
  done
done
python - <<PY
import json
for ln in open("$out"):
    try:
        d = json.loads(ln)
    except Exception as e:
        print("bad line", ln[:80]); continue
    r = d["run"]; print("abs=%d band=%3d N=%d residual %.4f ms" % (d["abs"], d["band"], r["N"], r["residual_ms"]))
PY
