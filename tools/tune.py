#!/usr/bin/env python
"""Kernel tuning sweep (run on the GPU box): times pass A / pass B in steady state for the
library named by $NKA_B200_LIB (default: the product build) and the grid overrides in
$NKA_GRID_PER_SM_A / _B.  Prints one JSON line.  Not part of the product."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from nka_b200 import NKA  # noqa: E402


def _peak():
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        return float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6554.9


PEAK = _peak()


def main():
    n = int(os.environ.get("TUNE_N", str(1 << 28)))
    m = int(os.environ.get("TUNE_M", "10"))
    steps = int(os.environ.get("TUNE_STEPS", "30"))
    lazy = os.environ.get("NKA_LAZY_LAST", "1") != "0"
    acc = NKA(n, m, 0.01)
    gen = torch.Generator(device="cuda").manual_seed(5)
    pool = [torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) - 0.5 for _ in range(m + 3)]
    k = 0
    for _ in range(m + 5):
        acc.accel_update(pool[k % len(pool)]); k += 1
    acc.synchronize()
    spans = os.environ.get("TUNE_SPANS", "1") != "0"     # 0: no per-kernel events (they break PDL overlap)
    acc.timing_enable(spans)
    acc.timing_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        acc.accel_update(pool[k % len(pool)]); k += 1
    e1.record()
    torch.cuda.synchronize()
    t = acc.timing_read()
    total = e0.elapsed_time(e1) / steps
    a = t["pass_a"]["ms"] / steps if spans else float("nan")
    b = t["pass_b"]["ms"] / steps if spans else float("nan")
    out = {
        "tag": os.environ.get("TUNE_TAG", "default"), "n": n, "m": m,
        "grid": acc.launch_geometry(), "ms_update": total, "ms_a": a, "ms_b": b,
        "ms_state": t["state"]["ms"] / steps, "ms_mat": t["materialise"]["ms"] / max(t["materialise"]["count"], 1),
        "lazy": lazy, "spans": spans, "pdl": os.environ.get("NKA_PDL", "1"),
        "tbs_a_actual": (m + (1 if lazy else 2)) * n * 8 / a / 1e9, "tbs_b_actual": (m + 4) * n * 8 / b / 1e9,
        "updates_per_s": 1e3 / total, "frac_roofline": (2 * m + 4) * n * 8 / (total * 1e-3) / 1e9 / PEAK,
        "hbm_gbs": (2 * m + 4) * n * 8 / (total * 1e-3) / 1e9, "peak_gbs": PEAK,
        "num_vec": acc.num_vec(),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
