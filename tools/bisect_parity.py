"""Debug aid: one parity scenario under the launch switches / library variants."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests'); "
        "import test_gpu_parity as T; T.test_scenarios_match_oracle(%r); print('ok')")
libs = {"product": None}
vdir = os.path.join(ROOT, "nka_b200", "lib", "variants")
if os.path.isdir(vdir):
    for f in sorted(os.listdir(vdir)):
        if f.endswith(".so"):
            libs[f] = os.path.join(vdir, f)
for name in sys.argv[1:] or ["iid_n1000_m10"]:
    for lib, path in libs.items():
        for pdl in ("1", "0"):
            env = dict(os.environ, NKA_PDL=pdl)
            if path:
                env["NKA_B200_LIB"] = path
            r = subprocess.run([sys.executable, "-c", code % (ROOT, ROOT, name)], env=env, capture_output=True, text=True, timeout=300)
            tail = (r.stdout.strip().splitlines() or [""])[-1] + " | " + (r.stderr.strip().splitlines() or [""])[-1]
            print("%-22s lib=%-28s pdl=%s -> rc=%d %s" % (name, lib, pdl, r.returncode, tail[:160]))
