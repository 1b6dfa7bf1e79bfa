#!/bin/bash
# ex_ssor_sweep2 build variants (run on the GPU box): bit-identity tests, then timing at 4096^2 and 8192^2.
# Usage: bash tools/ssor2_ab.sh <tag> <variant> [<variant> ...]   ("product" = the shipped library)
tag=$1; shift
out=gpurun_out/ssor2_ab_$tag.jsonl
: > $out
for lib in "$@"; do
  if [ $lib = product ]; then unset NKA_B200_LIB; else export NKA_B200_LIB=$PWD/nka_b200/lib/variants/libnka_b200_$lib.so; fi
  NKA_SSOR_KERNEL=2 timeout 600 python -m pytest tests/test_gpu_example.py -m gpu -q -x -k "sweep2 and (pc_ssor or full_size)" > gpurun_out/pytest_ssor2_${lib}_$tag.log 2>&1
  echo "$lib pytest rc=$? $(tail -1 gpurun_out/pytest_ssor2_${lib}_$tag.log)"
  for N in 4096 8192; do
    echo -n "{\"lib\": \"$lib\", \"run\": " >> $out
    NKA_SSOR_KERNEL=2 timeout 300 python tools/example_time.py $N 10 5 >> $out 2>> ${out%.jsonl}.err
    sed -i '$ s/$/}/' $out
  done
done
unset NKA_B200_LIB
python - <<PY
import json
for ln in open("$out"):
    try:
        d = json.loads(ln)
    except Exception as e:
        print("bad line", ln[:80]); continue
    r = d["run"]; print("%-16s N=%d ssor %.3f ms  residual %.4f  iter %.3f" % (d["lib"], r["N"], r["ssor_ms"], r["residual_ms"], r["ms_per_iter"]))
PY
