#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture, tuning sweep.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [tests] [bench] [ncu] [tune]
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_gpu_$tag.log
  tail -4 gpurun_out/pytest_gpu_$tag.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke_rc=$?" >> gpurun_out/smoke_$tag.log
  tail -2 gpurun_out/smoke_$tag.log ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench_rc=$?"
  cut -c1-400 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err ;;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
     python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-probe --no-sweep > gpurun_out/ncu_list_$tag.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:nka_pass -s 30 -c 4 -f -o gpurun_out/prof_$tag \
     python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-probe --no-sweep > gpurun_out/ncu_full_$tag.log 2>&1
  ls -la gpurun_out/prof_$tag.ncu-rep ;;
multi)
  ng=${NGPUS:-2}
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi_$tag.log 2>&1; echo "pytest_multi_rc=$?" >> gpurun_out/pytest_multi_$tag.log
  tail -3 gpurun_out/pytest_multi_$tag.log
  for g in 1 $ng; do
    if [ $g = 1 ]; then
      timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_g1_$tag.json 2> gpurun_out/bench_g1_$tag.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $g > gpurun_out/bench_g${g}_$tag.json 2> gpurun_out/bench_g${g}_$tag.err
    fi
    echo "bench g=$g rc=$?"; tail -2 gpurun_out/bench_g${g}_$tag.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_g${g}_$tag.json").read().strip().splitlines()[-1])
    print("g=$g value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "kern", {k:(v.get("avg_ms") if isinstance(v,dict) else v) for k,v in d["kernels"].items() if k!="geometry"})
except Exception as e: print("parse failed", e)
PY
  done ;;
sweep)
  : > gpurun_out/mvec_sweep_$tag.jsonl
  for m in 2 5 10 20; do
    TUNE_M=$m TUNE_TAG="mvec$m" timeout 300 python tools/tune.py >> gpurun_out/mvec_sweep_$tag.jsonl 2>> gpurun_out/mvec_sweep_$tag.err
  done
  cat gpurun_out/mvec_sweep_$tag.jsonl | cut -c1-600 ;;
example)
  for N in 4096; do
    timeout 600 python tools/example_time.py $N 20 5 >> gpurun_out/example_$tag.jsonl 2>> gpurun_out/example_$tag.err
  done
  cat gpurun_out/example_$tag.jsonl ;;
tune)
  : > gpurun_out/tune_$tag.jsonl
  for lib in ${TUNE_LIBS:-default t256_b2 t512_b1 t256_b1_st0 t512_b1_st0 t512_b1_ld0 t1024_b1}; do
    for g in ${TUNE_GRIDS:-0 2 4 8}; do
      if [ $lib = default ]; then unset NKA_B200_LIB; else export NKA_B200_LIB=$PWD/nka_b200/lib/variants/libnka_b200_$lib.so; fi
      if [ $g = 0 ]; then unset NKA_GRID_PER_SM_A NKA_GRID_PER_SM_B; else export NKA_GRID_PER_SM_A=$g NKA_GRID_PER_SM_B=$g; fi
      TUNE_TAG="$lib/g$g" timeout 120 python tools/tune.py >> gpurun_out/tune_$tag.jsonl 2>> gpurun_out/tune_$tag.err
    done
  done
  unset NKA_B200_LIB NKA_GRID_PER_SM_A NKA_GRID_PER_SM_B
  python - <<PY
import json
for ln in open("gpurun_out/tune_$tag.jsonl"):
    d=json.loads(ln); print("%-14s grid=%s upd=%.3fms A=%.3f (%.0f GB/s) B=%.3f (%.0f GB/s) state=%.3f frac=%.3f"%(d["tag"],(d["grid"]["grid_a"],d["grid"]["grid_b"]),d["ms_update"],d["ms_a"],d["tbs_a_actual"],d["ms_b"],d["tbs_b_actual"],d["ms_state"],d["frac_roofline"]))
PY
  ;;
esac
done
