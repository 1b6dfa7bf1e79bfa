#!/bin/bash
# Residual strip kernel, absolute bands: band height sweep on square grids and the 8-GPU slab shape (GPU box).
tag=$1
out=gpurun_out/residual_sweep_$tag.jsonl
: > $out
timeout 600 python -m pytest tests/test_gpu_example.py -m gpu -q -x -k "residual or golden or full_size or rectangular" > gpurun_out/pytest_residual_$tag.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_residual_$tag.log)"
for shape in "4096 4096" "8192 8192" "32768 4096" "2048 2048" "1024 1024"; do
  for band in ${BANDS:-24 32 48 64 auto}; do
    if [ $band = auto ]; then unset NKA_RES_BAND; else export NKA_RES_BAND=$band; fi
    timeout 120 python tools/residual_time.py $shape 20 >> $out 2>> ${out%.jsonl}.err
  done
done
unset NKA_RES_BAND
cat $out
