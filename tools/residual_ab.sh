#!/bin/bash
# Residual kernel variants at 4096^2 and 8192^2 (run on the GPU box).
out=${1:-gpurun_out/residual_ab.jsonl}
: > $out
for lib in product rs_mb2 rs_mb4 rs_pf8_mb2 rs_pf2_mb4; do
  if [ $lib = product ]; then unset NKA_B200_LIB; else export NKA_B200_LIB=$PWD/nka_b200/lib/variants/libnka_b200_$lib.so; fi
  for ipw in 2 4; do
    for N in 4096 8192; do
      echo -n "{\"lib\": \"$lib\", \"items_per_warp\": $ipw, \"run\": " >> $out
      NKA_RES_ITEMS_PER_WARP=$ipw timeout 300 python tools/example_time.py $N 10 5 >> $out 2>> ${out%.jsonl}.err
      sed -i '$ s/$/}/' $out
    done
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
    r = d["run"]; print("%-12s ipw=%d N=%d residual %.4f ms  ssor %.3f" % (d["lib"], d["items_per_warp"], r["N"], r["residual_ms"], r["ssor_ms"]))
PY
