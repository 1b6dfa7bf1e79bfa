#!/bin/bash
# compute-sanitizer over the device example's tests (run on the GPU box):
#   memcheck  -- every example test except the two 4096^2 ones
#   racecheck -- shared-memory hazards of ex_ssor_sweep2's producer / consumer rings on small grids
# Usage (under gpurun, from the repo root): bash tools/sanitize_ssor.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_example.py -m gpu -q -x \
  -k "not full_size" > gpurun_out/memcheck_example_$tag.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/memcheck_example_$tag.log
tail -4 gpurun_out/memcheck_example_$tag.log
NKA_SSOR_KERNEL=2 timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_example.py -m gpu -q -x \
  -k "pc_ssor_bit_identical and sweep2 and (33-70 or 96-40 or 300-300)" > gpurun_out/racecheck_ssor2_$tag.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/racecheck_ssor2_$tag.log
grep -c "Race reported\|hazard" gpurun_out/racecheck_ssor2_$tag.log
grep "Race reported\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/racecheck_ssor2_$tag.log | sort | uniq -c | sort -rn | head -20
