"""Long runs of the device example with both SSOR kernels (run on the GPU box): the residual
histories must agree bit for bit (every sweep is exact-order Gauss-Seidel in either kernel), and no
sweep may report a timed-out wait.  Usage: python tools/ssor_stress.py [N] [iters]  -> one JSON line"""
import json
import os
import subprocess
import sys

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 120

if len(sys.argv) > 3:                       # child: one kernel, prints the history as hex floats
    import torch
    sys.path.insert(0, ".")
    from nka_b200.example import System, Solver
    torch.cuda.set_device(0)
    sy = System(0.02, N, N, scaling=1)
    so = Solver(sy, nsweep=2, omega=1.4, mvec=5)
    out = so.solve(maxitr=iters)
    print(json.dumps({"iters": out["iters"], "rnorm": [float(x).hex() for x in out["rnorm"]]}))
    sys.exit(0)

hist = {}
for kv in ("2", "1"):
    env = dict(os.environ, NKA_SSOR_KERNEL=kv)
    r = subprocess.run([sys.executable, __file__, str(N), str(iters), "child"], env=env, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print(json.dumps({"N": N, "kernel": kv, "failed": r.stderr[-400:]}))
        sys.exit(1)
    hist[kv] = json.loads(r.stdout.strip().splitlines()[-1])
same = hist["1"]["rnorm"] == hist["2"]["rnorm"]
print(json.dumps({"N": N, "iters": hist["2"]["iters"], "sweeps_per_kernel": 4 * hist["2"]["iters"],
                  "histories_bit_identical": same, "rnorm_last": float.fromhex(hist["2"]["rnorm"][-1])}))
sys.exit(0 if same else 2)
