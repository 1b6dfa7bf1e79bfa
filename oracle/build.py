"""Build recipe for the parity checker (TEST INFRASTRUCTURE ONLY).

Two artefacts, both CPU code, neither ever linked or imported by nka_b200/:

* ``oracle/_build/libnka_oracle.so`` -- our own C restatement (nka_oracle.c).
* ``oracle/_ref/libnka_ref.so`` and ``oracle/_ref/nka_example_ref`` -- the
  reference's own C flavour, compiled *from the sources where they lie* under
  /root/reference/src-C (never copied into this repo).  Only possible in the
  build container; the GPU box has no /root/reference and uses the prebuilt
  files, which travel with the gpurun snapshot (git-ignored, not
  gpurun-ignored).

Only tests/, __graft_entry__ (build/smoke) and bench.py's cpu_baseline /
``--impl reference`` legs may call this module.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src-C"
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")

ORACLE_SO = os.path.join(BUILD_DIR, "libnka_oracle.so")
REF_SO = os.path.join(REF_DIR, "libnka_ref.so")
REF_EXAMPLE = os.path.join(REF_DIR, "nka_example_ref")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), proc.stdout, proc.stderr))


def build_oracle(force: bool = False) -> str:
    """Compile the restatement.  -ffp-contract=off keeps gcc from fusing a*b+c
    so the port rounds like the reference built for baseline x86-64."""
    src = os.path.join(HERE, "nka_oracle.c")
    if force or not _newer(ORACLE_SO, src, __file__):
        os.makedirs(BUILD_DIR, exist_ok=True)
        _run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
              "-o", ORACLE_SO, src, "-lm"])
    return ORACLE_SO


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "nonlinear_krylov_accelerator.c"))


def build_reference(force: bool = False) -> tuple[str | None, str | None]:
    """Compile the unmodified reference C flavour into oracle/_ref/.

    Flags follow the reference's Release configuration (-O3 -DNDEBUG,
    src-C/CMakeLists.txt).  Returns (library, example) paths; falls back to
    whatever prebuilt files exist when /root/reference is absent (GPU box)."""
    lib_c = os.path.join(REF_SRC, "nonlinear_krylov_accelerator.c")
    ex_c = os.path.join(REF_SRC, "nka_example.c")
    if reference_available():
        os.makedirs(REF_DIR, exist_ok=True)
        if force or not _newer(REF_SO, lib_c, __file__):
            _run(["gcc", "-O3", "-DNDEBUG", "-ffp-contract=off", "-fPIC", "-shared",
                  "-I", REF_SRC, "-o", REF_SO, lib_c, "-lm"])
        if force or not _newer(REF_EXAMPLE, lib_c, ex_c, __file__):
            _run(["gcc", "-O3", "-DNDEBUG", "-ffp-contract=off", "-I", REF_SRC,
                  "-o", REF_EXAMPLE, ex_c, lib_c, "-lm"])
    return (REF_SO if os.path.exists(REF_SO) else None,
            REF_EXAMPLE if os.path.exists(REF_EXAMPLE) else None)


def build_all(force: bool = False) -> dict:
    out = {"oracle": build_oracle(force)}
    lib, ex = build_reference(force)
    out["ref_lib"] = lib
    out["ref_example"] = ex
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(build_all(force=True), indent=1))
