"""Time the device example (BASELINE.json configs[1]): Picard iterations on an N x N grid.
Usage: python tools/example_time.py [N] [iters] [mvec]   -> one JSON line"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from nka_b200.example import System, Solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mvec = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
sy = System(0.02, N, N, scaling=1, stream=stream.cuda_stream)
so = Solver(sy, nsweep=2, omega=1.4, mvec=mvec)
so.solve(maxitr=3)                      # warm-up
sy.timing_enable(True)
if so.accel:
    so.accel.timing_enable(True)
    so.accel.timing_reset()
torch.cuda.synchronize()
t0 = time.perf_counter()
out = so.solve(maxitr=iters)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
kt = sy.timing_read()
line = {"N": N, "iters": out["iters"], "mvec": mvec, "ms_per_iter": 1e3 * dt / out["iters"],
        "ssor_ms": kt["pc_ssor"]["ms"] / max(kt["pc_ssor"]["count"], 1),
        "residual_ms": kt["residual"]["ms"] / max(kt["residual"]["count"], 1),
        "rnorm_last": out["rnorm"][-1]}
if so.accel:
    at = so.accel.timing_read()
    line["accel_ms"] = sum(v["ms"] for v in at.values()) / max(at["pass_b"]["count"], 1)
print(json.dumps(line))
