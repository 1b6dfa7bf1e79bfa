#!/bin/bash
# compute-sanitizer memcheck over what round 2 added (run on the GPU box): the residual strip kernel with absolute
# bands and its interior specialization (a grid large enough to have interior items), ex_ssor_sweep3, the pageable
# host path, the dp hook.  Usage (under gpurun, from the repo root): bash tools/sanitize_round2.sh <tag>
tag=$1
mkdir -p gpurun_out
log=gpurun_out/memcheck_round2_$tag.log
: > $log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_example.py -m gpu -q -x \
  -k "residual or (pc_ssor_bit_identical and sweep3) or rectangular" >> $log 2>&1
echo "memcheck example rc=$?" | tee -a $log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python - >> $log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from nka_b200.example import System, FIELD_U, FIELD_R
from oracle import api
# 700 x 420: strips and bands with interior items, edge items on all four sides
nx, ny = 700, 420
rng = np.random.default_rng(5)
u = rng.uniform(0.0, 0.3, (ny, nx))
sy = System(0.02, nx, ny, scaling=1)
sy.set(FIELD_U, u)
sy.residual()
orc = api.OracleSystem(nx, ny, 0.02, 1)
pad = np.zeros((ny + 2, nx + 2)); pad[1:-1, 1:-1] = u
assert np.array_equal(sy.get(FIELD_R), orc.residual(pad).reshape(ny, nx))
print("interior-path residual ok under memcheck")
PY
echo "memcheck interior rc=$?" | tee -a $log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "host_pointer_path_pipelined or dot_product_hook or drop_in" >> $log 2>&1
echo "memcheck host paths rc=$?" | tee -a $log
grep -E "ERROR SUMMARY|passed|failed|ok under" $log
