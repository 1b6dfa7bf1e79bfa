#!/bin/bash
# SSOR kernel comparison on the GPU box: parity of both kernels, per-strip trace, example timing,
# optional ncu source-level capture of ex_ssor_sweep2.
# Usage (under gpurun, from the repo root): bash tools/ssor_round.sh <tag> [ncu]
tag=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_example.py -m gpu -q -x > gpurun_out/pytest_example_$tag.log 2>&1
echo "pytest_rc=$?" >> gpurun_out/pytest_example_$tag.log
tail -5 gpurun_out/pytest_example_$tag.log
for kv in ${KERNELS:-1 2}; do
  NKA_SSOR_KERNEL=$kv timeout 120 python tools/ssor_trace.py 4096 > gpurun_out/ssor_trace_k${kv}_$tag.txt 2>&1
  echo "kernel $kv trace rc=$?"; head -c 1200 gpurun_out/ssor_trace_k${kv}_$tag.txt; echo
  for N in ${SIZES:-1024 4096 8192}; do
    NKA_SSOR_KERNEL=$kv timeout 300 python tools/example_time.py $N 10 5 >> gpurun_out/example_k${kv}_$tag.jsonl 2>> gpurun_out/example_k${kv}_$tag.err
  done
  cat gpurun_out/example_k${kv}_$tag.jsonl
done
if [ "$2" = ncu ]; then
  NKA_SSOR_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ex_ssor_sweep2 -s 8 -c 2 -f \
    -o gpurun_out/prof_ssor2_$tag python tools/example_time.py 4096 3 5 > gpurun_out/ncu_ssor2_$tag.log 2>&1
  ls -la gpurun_out/prof_ssor2_$tag.ncu-rep
fi
