"""Per-strip timeline of one SSOR sweep (tuning aid; run on the GPU box).
Usage: python tools/ssor_trace.py [N]"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nka_b200 import _lib  # noqa: E402
from nka_b200.example import System  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
sy = System(0.02, N, N, scaling=1)
sy.residual()
sy.pc_ssor(1, 1.4)            # warm-up
ns = lib.nka_system_ssor_trace(sy._h, 1, None)
sy.pc_ssor(1, 1.4)            # forward + backward: the trace holds the backward sweep (last writer)
buf = np.zeros(ns * 4 + 256, dtype=np.uint64)
lib.nka_system_ssor_trace(sy._h, 0, buf.ctypes.data)
tr = buf[:ns * 4].reshape(ns, 4).astype(np.int64)
steps_t = buf[ns * 4:].astype(np.int64)
t0 = tr[:, 0].min()
start, first, end, waits = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2] - t0, tr[:, 3]
order = np.argsort(first + (first <= 0) * 10**15)
steps = N + (63 if int(__import__("os").environ.get("NKA_SSOR_KERNEL", "2")) == 3 else 31)
out = {"N": N, "nstrips": int(ns), "sweep_us": float((end.max()) / 1e3),
       "strip_walk_us_median": float(np.median(end - np.maximum(first, start)) / 1e3),
       "ns_per_step_in_walk_median": float(np.median(end - np.maximum(first, start)) / steps),
       "first_strip_walk_us": float((end - start).min() / 1e3),
       "lag_between_strips_us_median": float(np.median(np.diff(np.sort(first[first > 0]))) / 1e3) if (first > 0).sum() > 2 else None,
       "sync_waits_per_strip_median": float(np.median(waits)), "sync_waits_max": int(waits.max())}
print(json.dumps(out))
d = np.diff(steps_t)
print('mid-strip per-step ns, steps 0..255:', d.tolist())
print('mid strip: start->step0 us', (steps_t[0] - t0 - start[ns // 2]) / 1e3)
