#!/bin/bash
# Residual strip kernel: L2 prefetch distance variants (GPU box).
tag=$1; shift
out=gpurun_out/residual_l2_$tag.jsonl
: > $out
for lib in "$@"; do
  if [ $lib = product ]; then unset NKA_B200_LIB; else export NKA_B200_LIB=$PWD/nka_b200/lib/variants/libnka_b200_$lib.so; fi
  timeout 300 python -m pytest tests/test_gpu_example.py -m gpu -q -x -k "residual" > gpurun_out/pytest_res_${lib}_$tag.log 2>&1
  echo "$lib pytest rc=$? $(tail -1 gpurun_out/pytest_res_${lib}_$tag.log)"
  for shape in "4096 4096" "8192 8192" "32768 4096"; do
    echo -n "{\"lib\": \"$lib\", \"run\": " >> $out
    timeout 120 python tools/residual_time.py $shape 20 >> $out 2>> ${out%.jsonl}.err
    sed -i '$ s/$/}/' $out
  done
done
unset NKA_B200_LIB
python - <<PY
import json
for ln in open("$out"):
    try:
        d = json.loads(ln)
    except Exception:
        print("bad line", ln[:80]); continue
    r = d["run"]; print("%-10s %5d x %5d  %.4f ms  %.0f GB/s" % (d["lib"], r["nx"], r["ny"], r["residual_ms"], r["gbs_algorithmic"]))
PY
