"""Time the device example on row slabs (BASELINE.json configs[3]).  Launch with torchrun, one process
per GPU (or plain python for 1 GPU):
    python -m torch.distributed.run --nproc-per-node G tools/example_time_dist.py [N] [iters] [mvec]
Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from nka_b200.example import System, Solver, distributed_system  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mvec = int(sys.argv[3]) if len(sys.argv) > 3 else 5
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
stream = torch.cuda.Stream()
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sy = distributed_system(0.02, N, N, scaling=1, device=local, stream=stream.cuda_stream)
else:
    sy = System(0.02, N, N, scaling=1, device=local, stream=stream.cuda_stream)
so = Solver(sy, nsweep=2, omega=1.4, mvec=mvec)
warm = so.solve(maxitr=2)
sy.timing_enable(True)
if so.accel:
    so.accel.timing_enable(True)
    so.accel.timing_reset()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = so.solve(maxitr=iters)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
kt = sy.timing_read()
line = {"N": N, "gpus": world, "rows_per_gpu": sy.ny, "iters": int(out["iters"]), "mvec": mvec,
        "ms_per_iter": 1e3 * dt / out["iters"],
        "ssor_ms": kt["pc_ssor"]["ms"] / max(kt["pc_ssor"]["count"], 1),
        "residual_ms": kt["residual"]["ms"] / max(kt["residual"]["count"], 1),
        "rnorm_first": [float(x) for x in warm["rnorm"]] + [float(x) for x in out["rnorm"][1:6]],
        "rnorm_last": float(out["rnorm"][-1]),
        "comm_mode": so.accel.comm_mode() if so.accel else None}
if so.accel:
    at = so.accel.timing_read()
    line["accel_ms"] = sum(v["ms"] for v in at.values()) / max(at["pass_b"]["count"], 1)
if world > 1:
    t = torch.tensor([line["ms_per_iter"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    line["ms_per_iter"] = float(t.item())
if rank == 0:
    print(json.dumps(line))
so.delete(); sy.delete()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
