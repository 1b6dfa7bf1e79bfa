"""Phase timeline of pass A (tuning aid; needs the NKA_TRACE variant: python -c "from nka_b200 import build;
build.build_variant('trace', {'NKA_TRACE': 1})" then NKA_B200_LIB=nka_b200/lib/variants/libnka_b200_trace.so).
Usage: python tools/pass_a_trace.py [n] [mvec]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nka_b200 import NKA, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 25
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = _lib.load()
lib.nka_debug_trace.argtypes = [C.c_void_p]
acc = NKA(n, m, 0.01)
gen = torch.Generator(device="cuda").manual_seed(5)
pool = [torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) - 0.5 for _ in range(m + 3)]
buf = (C.c_ulonglong * 8)()
k = 0
for _ in range(m + 5):
    acc.accel_update(pool[k % len(pool)]); k += 1
lib.nka_debug_trace(buf)
rows = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc.accel_update(pool[k % len(pool)]); k += 1
    e1.record()
    lib.nka_debug_trace(buf)
    t = np.array(list(buf), dtype=np.int64)
    rows.append([e0.elapsed_time(e1) * 1e3] + [(t[i] - t[0]) / 1e3 for i in range(1, 6)])
r = np.median(np.array(rows), axis=0)
print(json.dumps({"n": n, "mvec": m, "grid": acc.launch_geometry(), "update_us_events": r[0],
                  "us_since_first_cta_start": {"last_loop_end": r[1], "ticket_won": r[2], "rows_folded": r[3],
                                               "exchange_done": r[4], "state_committed": r[5]}}))
