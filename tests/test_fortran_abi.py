"""The Fortran layers cannot be compiled here (no Fortran compiler in the image), so the
binding surface is checked from the C side: every `bind(C, name=...)` in
nka_b200/fortran/nka_b200_c.F90 must be exported by the library, declared in include/*.h with
the same number of arguments, and pass scalars by value exactly where the header does."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F90 = os.path.join(ROOT, "nka_b200", "fortran", "nka_b200_c.F90")
F90_EXAMPLE = os.path.join(ROOT, "nka_b200", "fortran", "nka_example_c.F90")


def _fortran_interfaces():
    text = open(F90).read() + open(F90_EXAMPLE).read()
    text = text.replace("&\n", " ")
    out = {}
    pat = re.compile(r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)(.*?)end (?:function|subroutine)",
                     re.S | re.I)
    for m in pat.finditer(text):
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        body = m.group(4)
        by_value = set()
        for line in body.splitlines():
            if "value" in line.lower() and "::" in line:
                by_value.update(n.strip() for n in line.split("::")[1].split(","))
        out[m.group(3)] = {"args": args, "by_value": by_value}
    return out


def _c_declarations():
    decls = {}
    for fn in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(nka_\w+)\s*\(([^;{]*?)\)\s*;", text, re.S):
            params = m.group(2).strip()
            params = re.sub(r"\([^()]*\)", "", params)        # drop nested (fn-pointer) parens
            n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
            decls[m.group(1)] = {"n": n, "params": [p.strip() for p in m.group(2).split(",")]}
    return decls


def test_every_fortran_binding_is_exported_and_declared():
    from nka_b200 import build
    path = build.build_library()
    exported = set(subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True,
                                  check=True).stdout.split())
    ifaces = _fortran_interfaces()
    decls = _c_declarations()
    assert len(ifaces) >= 55
    for name, info in ifaces.items():
        assert name in exported, name
        assert name in decls, name
        if name == "nka_init":
            continue
        assert len(info["args"]) == decls[name]["n"], (name, info["args"], decls[name])


def test_scalars_by_value_arrays_by_address():
    """A C `double`/`int`/`size_t`/handle parameter must be `value` in Fortran; a `double *` array
    or the 128-byte id must not be."""
    ifaces = _fortran_interfaces()
    decls = _c_declarations()
    for name, info in ifaces.items():
        for farg, cparam in zip(info["args"], decls[name]["params"]):
            is_array = ("double *" in cparam and "NKA" not in cparam) or "id128" in cparam
            if name in ("nka_accel_update_dev",) and "f_dev" in farg:
                is_array = False          # device address travels by value in a type(c_ptr)
            assert (farg in info["by_value"]) == (not is_array), (name, farg, cparam)


def test_fortran_modules_keep_reference_names():
    """Module, type and procedure names of the three reference flavours are kept."""
    f95 = open(os.path.join(ROOT, "nka_b200", "fortran", "F95", "nka_type.F90")).read()
    for name in ("nka_init", "nka_delete", "nka_set_vec_tol", "nka_defined", "nka_vec_len", "nka_num_vec",
                 "nka_max_vec", "nka_vec_tol", "nka_real_kind", "nka_accel_update", "nka_relax", "nka_restart"):
        assert re.search(r"public ::.*\b%s\b" % name, f95), name       # src-F95/nka_type.F90:205-207
    assert "module nka_type" in f95 and "type, public :: nka" in f95
    f08 = open(os.path.join(ROOT, "nka_b200", "fortran", "F08", "nka_type.F90")).read()
    for name in ("init", "set_vec_tol", "set_dot_prod", "vec_len", "num_vec", "max_vec", "vec_tol",
                 "accel_update", "relax", "restart", "defined"):                 # src-F08/nka_type.F90:169-181
        assert re.search(r"(procedure|generic)\s*(,\s*private)?\s*::\s*%s\b" % name, f08), name
    vec = open(os.path.join(ROOT, "nka_b200", "fortran", "F08-vector", "gpu_vector_type.F90")).read()
    for name in ("clone1", "clone2", "copy_", "setval", "scale", "update1_", "update2_", "update3_", "update4_",
                 "dot_", "norm2"):                                               # vector_class.F90:93-108
        assert re.search(r"procedure :: %s\b" % name, vec), name
    assert "type, extends(vector), public :: gpu_vector" in vec
