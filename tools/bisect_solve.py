"""Debug aid: the 50x50 accelerated solve under the kernel-selection switches (iters should be 26)."""
import os, subprocess, sys, json
code = r'''
import sys; sys.path.insert(0, ".")
from nka_b200.example import System, Solver
sy = System(0.02, 50, 50, scaling=0)
so = Solver(sy, nsweep=2, omega=1.4, mvec=5, vtol=0.01)
out = so.solve(record_nvec=True)
print(out["iters"], list(out["rnorm"][:4]), list(out["nvec"][:8]))
'''
for pdl in ("1", "0"):
    for rk in ("2", "1"):
        for sk in ("2", "1"):
            env = dict(os.environ, NKA_PDL=pdl, NKA_RESIDUAL_KERNEL=rk, NKA_SSOR_KERNEL=sk)
            r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
            print("pdl=%s residual=%s ssor=%s ->" % (pdl, rk, sk), r.stdout.strip()[:300], r.stderr.strip()[-300:])
