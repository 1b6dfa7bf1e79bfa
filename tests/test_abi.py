"""The C-ABI library builds for sm_100a, loads, and exports every symbol that
include/*.h declares.  No compute calls: this runs without a GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared_functions():
    names = set()
    for fn in os.listdir(INC):
        text = open(os.path.join(INC, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(nka_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    return names


def test_library_builds_and_exports_every_declared_symbol():
    from nka_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    exported = set(subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True,
                                  check=True).stdout.split())
    declared = _declared_functions()
    assert len(declared) >= 25
    missing = sorted(n for n in declared if n not in exported)
    assert not missing, missing
    # the binding table names exactly the declared functions
    assert set(_lib.SYMBOLS) == declared
    lib = _lib.load()
    assert lib.nka_b200_version().startswith(b"nka_b200")


def test_reference_header_signatures_are_kept():
    """The nine functions of src-C/nonlinear_krylov_accelerator.h:3-12, verbatim shapes."""
    text = open(os.path.join(INC, "nonlinear_krylov_accelerator.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    flat = " ".join(text.split())
    for decl in [
        "typedef struct nka_state *NKA;",
        "extern NKA nka_init (int vlen, int mvec, double vtol, double (*dp)(int, double *, double *));",
        "extern void nka_delete (NKA);",
        "extern void nka_accel_update (NKA, double *f);",
        "extern void nka_restart (NKA);",
        "extern void nka_relax (NKA);",
        "extern int nka_num_vec (NKA);",
        "extern int nka_max_vec (NKA);",
        "extern int nka_vec_len (NKA);",
        "extern double nka_vec_tol (NKA);",
    ]:
        assert decl in flat, decl


def test_sass_is_sm100a_with_128bit_loads():
    """cuobjdump: the library carries sm_100a code only and the streaming kernels use LDG.E.128."""
    from nka_b200 import build
    path = build.build_library()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_python_binding_validates_like_the_reference_asserts():
    """src-F08/nka_type.F90:190-191,205: mvec > 0, vlen >= 0, vtol > 0 -- raised before any CUDA call."""
    from nka_b200 import NKA
    acc = NKA()
    for bad in [dict(vlen=10, mvec=0), dict(vlen=10, mvec=-1), dict(vlen=-1, mvec=3),
                dict(vlen=10, mvec=3, vtol=0.0), dict(vlen=10, mvec=33)]:
        with pytest.raises(ValueError):
            acc.init(**bad)
    assert not acc.defined()


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the library refuses to create an accelerator (abort with a message)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    code = ("import sys; sys.path.insert(0, %r); from nka_b200 import NKA; NKA(16, 2)" % ROOT)
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr
