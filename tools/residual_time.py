"""Time the fused residual kernel alone on an NX x NY grid (event-timed inside the library).
Usage: python tools/residual_time.py NX NY [reps]   -> one JSON line"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nka_b200.example import System, FIELD_U  # noqa: E402

nx, ny = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
torch.cuda.set_device(0)
sy = System(0.02, nx, ny, scaling=1)
for _ in range(3):
    sy.residual(True)
sy.timing_enable(True)
for _ in range(reps):
    sy.residual(True)
kt = sy.timing_read()
ms = kt["residual"]["ms"] / max(kt["residual"]["count"], 1)
print(json.dumps({"nx": nx, "ny": ny, "residual_ms": ms, "gbs_algorithmic": 7 * 8 * nx * ny / ms / 1e6,
                  "band": os.environ.get("NKA_RES_BAND", "auto"), "abs": os.environ.get("NKA_RES_ABS_BANDS", "1"),
                  "ipw": os.environ.get("NKA_RES_ITEMS_PER_WARP", "4")}))
