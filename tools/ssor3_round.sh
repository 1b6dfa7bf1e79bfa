#!/bin/bash
# ex_ssor_sweep3 against ex_ssor_sweep2 on the GPU box: parity of all kernels, per-strip trace, example timing.
# Usage (under gpurun, from the repo root): bash tools/ssor3_round.sh <tag> [ncu]
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_example.py -m gpu -q -x -k "ssor or example or full_size" > gpurun_out/pytest_example_$tag.log 2>&1
echo "pytest_rc=$?" >> gpurun_out/pytest_example_$tag.log
tail -5 gpurun_out/pytest_example_$tag.log
for kv in ${KERNELS:-2 3}; do
  NKA_SSOR_KERNEL=$kv timeout 120 python tools/ssor_trace.py 4096 > gpurun_out/ssor_trace_k${kv}_$tag.txt 2>&1
  echo "kernel $kv trace rc=$?"; head -c 900 gpurun_out/ssor_trace_k${kv}_$tag.txt; echo
  for N in ${SIZES:-1024 4096 8192}; do
    NKA_SSOR_KERNEL=$kv timeout 300 python tools/example_time.py $N 10 5 >> gpurun_out/example_k${kv}_$tag.jsonl 2>> gpurun_out/example_k${kv}_$tag.err
  done
  cat gpurun_out/example_k${kv}_$tag.jsonl
done
if [ "$2" = ncu ]; then
  NKA_SSOR_KERNEL=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ex_ssor_sweep3 -s 8 -c 2 -f \
    -o gpurun_out/prof_ssor3_$tag python tools/example_time.py 4096 3 5 > gpurun_out/ncu_ssor3_$tag.log 2>&1
  ls -la gpurun_out/prof_ssor3_$tag.ncu-rep
fi
