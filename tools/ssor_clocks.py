"""Where one consumer step of ex_ssor_sweep2 spends its cycles (tuning aid; needs a library built with -DEX2_CLOCKS,
e.g. nka_b200.build.build_variant("k2_clocks", {"EX2_CLOCKS": 1}), NKA_B200_LIB pointing at it; run on the GPU box).
Stamps per step of the middle strip, steps 1024..1087: top of the step, upstream value selected, numerator x formed,
result zc formed.  The stamps themselves perturb the schedule (volatile asm), so the total is larger than in the
product build; the split is what is of interest."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from nka_b200 import _lib  # noqa: E402
from nka_b200.example import System  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
sy = System(0.02, N, N, scaling=1)
sy.residual()
sy.pc_ssor(1, 1.4)
ns = lib.nka_system_ssor_trace(sy._h, 1, None)
sy.pc_ssor(1, 1.4)
buf = np.zeros(ns * 4 + 256, dtype=np.uint64)
lib.nka_system_ssor_trace(sy._h, 0, buf.ctypes.data)
st = (buf[ns * 4:].astype(np.int64) & 0xFFFFFFFF).reshape(64, 4)
d = lambda a, b: ((a - b) & 0xFFFFFFFF)
seg = {"top->zh (shuffle, select)": d(st[:, 1], st[:, 0]), "zh->x (mul, 4 adds, mul)": d(st[:, 2], st[:, 1]),
       "x->zc (division, add, branch)": d(st[:, 3], st[:, 2]), "zc->next top (stores, counters, loads, hand-over)": d(st[1:, 0], st[:-1, 3]),
       "whole step (top->top)": d(st[1:, 0], st[:-1, 0])}
out = {k: {"median": float(np.median(v)), "p10": float(np.percentile(v, 10)), "p90": float(np.percentile(v, 90))} for k, v in seg.items()}
print(json.dumps({"N": N, "cycles": out}))
print("per step (top->top), 63 steps:", d(st[1:, 0], st[:-1, 0]).tolist())
